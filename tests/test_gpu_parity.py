"""GPU parity tests proper: the CUDA path, called through the C-ABI, against the oracle and the
golden fixtures.  Bit-exact for both interpolations (north_star allows 1 LSB for Lanczos4; the
tolerance used here is 0), blur scores from exact integer sums (checked to 1e-12 relative against
the reference's fp64 values; north_star's bound is 1e-4)."""
import hashlib
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, load_json, noise_frame, smooth_frame, unhex

pytestmark = pytest.mark.gpu

LANCZOS_TOL_LSB = 0
SCORE_RTOL = 1e-12


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def assert_same(got, want, what, tol=0):
    if tol == 0:
        if not np.array_equal(got, want):
            bad = np.argwhere(got != want)
            raise AssertionError(f"{what}: {len(bad)} differing bytes, first at {bad[0].tolist()}: "
                                 f"got {got[tuple(bad[0])]} want {want[tuple(bad[0])]}")
    else:
        d = np.abs(got.astype(np.int16) - want.astype(np.int16)).max()
        assert d <= tol, f"{what}: max |diff| {d} > {tol}"


def rotations(x360, views, active=None):
    return x360.rotation_stack(views, active)[1]


# ---- golden fixtures (outputs of the real reference) -----------------------------------------
def test_small_golden_cases(x360):
    z = np.load(GOLDEN + "/small_cases.npz")
    meta = load_json("small_cases.json")
    src = z["src"]
    for k, m in enumerate(meta):
        R = np.array([unhex(v) for v in m["R"]]).reshape(1, 3, 3)
        for mode, key in (("linear", "lin"), ("lanczos", "lan")):
            with x360.Extractor(src.shape[0], src.shape[1], m["d"], m["fov"], R, mode) as ex:
                out, scores = ex.extract(src[None], want_scores=True)
                assert_same(out[0, 0], z[f"{key}{k}"], f"case {k} {mode}")
                assert scores[0, 0] == pytest.approx(unhex(m[f"{key}_score"]), rel=SCORE_RTOL)
                mx, my = ex.make_maps(0)
                assert np.array_equal(mx, z[f"mx{k}"]) and np.array_equal(my, z[f"my{k}"]), f"maps case {k}"


@pytest.mark.parametrize("case", ["cube,6,0,90", "ring,8,-20,90", "fibonacci,20,-20,110"])
def test_known_answer_hashes(x360, case):
    """512x1024 noise, d=256, sha256 over the views in generate_views order."""
    k = load_json("kat_512x1024.json")
    g = k["cases"][case]
    layout, n, pitch, fov = case.split(",")
    src = np.random.default_rng(12345).integers(0, 256, (512, 1024, 3), dtype=np.uint8)
    views = x360.GeometryProcessor.generate_views(int(n), pitch_offset=int(pitch), layout_mode=layout)
    R = rotations(x360, views)
    assert [[float(v).hex() for v in r.reshape(-1)] for r in R] == g["R"]
    for mode, key in (("linear", "linear"), ("lanczos", "lanczos")):
        with x360.Extractor(512, 1024, 256, int(fov), R, mode) as ex:
            out, scores = ex.extract(src[None], want_scores=True)
        assert [sha(out[0, v]) for v in range(len(views))] == g[f"{key}_view_sha"]
        assert sha(*[out[0, v] for v in range(len(views))]) == g[f"{key}_sha"]
        want = np.array([unhex(s) for s in g[f"{key}_scores"]])
        np.testing.assert_allclose(scores[0], want, rtol=SCORE_RTOL)


# ---- seeded sweeps against the oracle ----------------------------------------------------------
SWEEP = [
    # (H, W, d, fov, layout, n, pitch, F)
    (256, 512, 96, 90, "cube", 6, 0, 3),          # poles, seam, multi-frame double buffering
    (256, 512, 96, 90, "cube", 6, -45, 2),
    (192, 384, 80, 60, "ring", 8, -20, 2),
    (192, 384, 64, 120, "ring", 3, 45, 1),
    (200, 400, 95, 110, "fibonacci", 7, -20, 2),  # odd d: partial tiles, unaligned rows
    (161, 322, 50, 90, "fibonacci", 5, 20, 2),    # W % 4 != 0: generic gather everywhere
    (128, 256, 33, 100, "ring", 2, -90, 1),       # looking straight down
    (96, 192, 130, 75, "ring", 4, 0, 1),          # magnification (d larger than the source arc)
]


# path: 0 = auto (the tiled TMA kernel, both interpolations), 1 = direct gathers, 2 = the earlier staged kernel (bilinear)
# score variants of the tiled path: the split pass (default), fused on the raw gather, fused on the pair-record planes
@pytest.mark.parametrize("mode,flags", [("linear", 2), ("linear", 3), ("lanczos", 2)], ids=["linear-fused-raw", "linear-fused-planes", "lanczos-fused"])
@pytest.mark.parametrize("cfg", SWEEP)
def test_sweep_fused_score_variants(x360, oracle, cfg, mode, flags):
    H, W, d, fov, layout, n, pitch, F = cfg
    interp = 1 if mode == "lanczos" else 0
    frames = np.stack([noise_frame(100 + i, H, W) for i in range(F)])
    R = rotations(x360, x360.GeometryProcessor.generate_views(n, pitch_offset=pitch, layout_mode=layout))
    want, ws, ws2 = oracle.extract(frames, R, d, fov, interp)
    with x360.Extractor(H, W, d, fov, R, mode, flags=flags) as ex:
        got, s, s2 = ex.extract_sums(frames)
    assert_same(got, want, f"{cfg} {mode} flags {flags}", LANCZOS_TOL_LSB if interp else 0)
    assert np.array_equal(s, ws) and np.array_equal(s2, ws2), "integer Laplacian sums"


@pytest.mark.parametrize("mode,interp,path", [("linear", 0, 0), ("linear", 0, 1), ("linear", 0, 2), ("lanczos", 1, 0), ("lanczos", 1, 1)])
@pytest.mark.parametrize("cfg", SWEEP)
def test_sweep_against_oracle(x360, oracle, cfg, mode, interp, path):
    H, W, d, fov, layout, n, pitch, F = cfg
    frames = np.stack([noise_frame(100 + i, H, W) for i in range(F)])
    views = x360.GeometryProcessor.generate_views(n, pitch_offset=pitch, layout_mode=layout)
    R = rotations(x360, views)
    want, ws, ws2 = oracle.extract(frames, R, d, fov, interp)
    try:
        ex = x360.Extractor(H, W, d, fov, R, mode, path=path)
    except x360.X360Error as e:
        assert path in (1, 2) and e.code == -4, e          # the comparison kernels exist only in -DX360_LAB builds
        pytest.skip("comparison path not compiled into the default build")
    with ex:
        assert ex.info()["path"] == (1 if path == 1 else (2 if path == 2 else 3))
        got, s, s2 = ex.extract_sums(frames)
        modes = ex.tile_modes()
        got_ns, none = ex.extract(frames, want_scores=False)      # the no-score kernel variant
    assert none is None
    if interp == 0 and path in (0, 2):
        # tiles are 32 wide, 32 or 16 high; `modes` is of the extraction launch (the score is a separate pass by default)
        tile_hs = x360.plan_preview(H, W, d, fov, R)["tile_h_of_view"] if path == 0 else [32] * len(views)
        assert sum(modes.values()) == sum(((d + 31) // 32) * ((d + th - 1) // th) for th in tile_hs)
        scale = (W / (2 * np.pi)) / ((d / 2) / np.tan(np.radians(fov) / 2))     # source px per output px at the centre
        if W % 16 == 0 and scale < 1.5:                  # footprints too large for a profitable staging capacity go direct
            assert modes["staged"] + modes["staged_wide"] + modes["staged_seam"] > 0, modes
    assert_same(got, want, f"{cfg} {mode}", LANCZOS_TOL_LSB if interp else 0)
    assert_same(got_ns, want, f"{cfg} {mode} (no scores)", LANCZOS_TOL_LSB if interp else 0)
    assert np.array_equal(s, ws) and np.array_equal(s2, ws2), "integer Laplacian sums"


# ---- the raw-gather variant (bilinear, column-only map_x, no fused score) and the pair-record planes it replaced --------
RAW_CASES = [
    # (H, W, d, fov, n, F)
    (512, 1024, 256, 90, 8, 7),        # 1.27 source px per output px; 7 frames: the three raw buffers wrap around twice
    (480, 960, 250, 90, 6, 5),         # partial tiles, output pitch not a multiple of 16 B: byte stores instead of the TMA store
    (720, 1440, 512, 100, 4, 4),
    (360, 720, 288, 60, 3, 3),         # magnification: records shared by neighbouring lanes
    (1440, 2880, 800, 90, 8, 5),       # cfg2 at half size
    (300, 600, 128, 90, 4, 3),         # frame pitch not a multiple of 16 B: no TMA loads, every tile takes the direct gathers
]


@pytest.mark.parametrize("no_raw", ["0", "1"], ids=["raw", "planes"])
@pytest.mark.parametrize("cfg", RAW_CASES)
def test_raw_gather_ring_plans(x360, oracle, cfg, no_raw):
    H, W, d, fov, n, F = cfg
    flags = x360.FLAG_PLANES if no_raw == "1" else 0
    frames = np.stack([noise_frame(300 + i, H, W) for i in range(F)])
    views = x360.GeometryProcessor.generate_views(n, 0, "ring")
    R = rotations(x360, views)
    assert x360.column_only_map_x(R)
    want, _, _ = oracle.extract(frames, R, d, fov, 0, want_scores=False)
    with x360.Extractor(H, W, d, fov, R, "linear", max_batch=F, flags=flags) as ex:
        got, none = ex.extract(frames, want_scores=False)
        modes = ex.tile_modes()
        again, _ = ex.extract(frames, want_scores=False)
    assert none is None
    assert_same(got, want, f"{cfg} no_raw={no_raw}")
    assert np.array_equal(got, again)
    staged = modes["staged"] + modes["staged_wide"] + modes["staged_seam"]
    if (W * 3) % 16 == 0:
        assert staged > 0.5 * sum(modes.values()), modes
        if no_raw == "0" and W * 3 >= 2048 and n % 2 == 0:
            assert modes["staged_seam"] > 0, modes           # an even ring has a view at yaw 180: the seam runs through it, staged through the window
    else:
        assert staged == 0


@pytest.mark.parametrize("cfg", [(256, 512, 96, 90, "cube", 6, 0, 3), (192, 384, 80, 60, "ring", 8, -20, 2), (512, 1024, 256, 90, "ring", 8, 0, 5)])
def test_raw_gather_with_fused_score_opt_in(x360, oracle, cfg):
    """FLAG_FUSED_SCORE: the raw gather with the fused score (double-buffered gray tile and per-warp sums)."""
    H, W, d, fov, layout, n, pitch, F = cfg
    frames = np.stack([noise_frame(500 + i, H, W) for i in range(F)])
    R = rotations(x360, x360.GeometryProcessor.generate_views(n, pitch_offset=pitch, layout_mode=layout))
    want, ws, ws2 = oracle.extract(frames, R, d, fov, 0)
    with x360.Extractor(H, W, d, fov, R, "linear", max_batch=F, flags=x360.FLAG_FUSED_SCORE) as ex:
        got, s, s2 = ex.extract_sums(frames)
    assert_same(got, want, f"{cfg} raw + scores")
    assert np.array_equal(s, ws) and np.array_equal(s2, ws2), "integer Laplacian sums"


@pytest.mark.parametrize("cfg", [(256, 512, 96, 90, "cube", 6, 0, 3), (200, 400, 95, 110, "fibonacci", 7, -20, 2), (512, 1024, 260, 90, "ring", 8, 0, 2),
                                 (128, 256, 33, 100, "ring", 2, -90, 1)])
def test_split_score_pass_matches_the_fused_one(x360, oracle, cfg):
    """FLAG_SPLIT_SCORE: the Laplacian sums from one streaming pass over the finished views instead of the fused reduction."""
    H, W, d, fov, layout, n, pitch, F = cfg
    frames = np.stack([noise_frame(600 + i, H, W) for i in range(F)])
    R = rotations(x360, x360.GeometryProcessor.generate_views(n, pitch_offset=pitch, layout_mode=layout))
    want, ws, ws2 = oracle.extract(frames, R, d, fov, 0)
    with x360.Extractor(H, W, d, fov, R, "linear", max_batch=F, flags=x360.FLAG_SPLIT_SCORE) as ex:
        got, s, s2 = ex.extract_sums(frames)
    assert_same(got, want, f"{cfg} split score")
    assert np.array_equal(s, ws) and np.array_equal(s2, ws2), "integer Laplacian sums"


@pytest.mark.parametrize("cfg", [(510, 1020, 228, 90.0, 10, 45, True), (97, 972, 70, 80.0, 11, 50, True), (256, 512, 128, 90.0, 8, 45, False),
                                 (256, 512, 96, 60.0, 6, 60, False)])
def test_views_with_a_pixel_on_the_pole_do_not_share_a_map(x360, oracle, cfg):
    """pitch + fov / 2 = 90 deg and an even output size: the top (bottom) centre pixel looks exactly at a pole, where map_x is
    decided by each view's own rounding residues (found by scripts/fuzz_parity.py).  Such views evaluate their own map."""
    H, W, d, fov, n, pitch, scores = cfg
    frames = np.stack([noise_frame(800 + i, H, W) for i in range(3)])
    views = x360.GeometryProcessor.generate_views(n, pitch_offset=pitch, layout_mode="ring")
    R = rotations(x360, views)[:8]
    assert x360.plan_preview(H, W, d, fov, R)["n_units"] == len(R)
    want, ws, ws2 = oracle.extract(frames, R, d, fov, 0, want_scores=scores)
    with x360.Extractor(H, W, d, fov, R, "linear", max_batch=3) as ex:
        if scores:
            got, s, s2 = ex.extract_sums(frames)
            assert np.array_equal(s, ws) and np.array_equal(s2, ws2)
        else:
            got, _ = ex.extract(frames, want_scores=False)
    assert_same(got, want, f"{cfg} pole pixel")


def test_roll_and_active_camera_subset(x360, oracle):
    views = x360.GeometryProcessor.generate_views(12, 0, "ring")
    names, R = x360.rotation_stack(views, active_cameras=[0, 3, 6, 9])
    frame = noise_frame(8, 256, 512)
    for mode, interp in (("linear", 0), ("lanczos", 1)):
        want, _, _ = oracle.extract(frame[None], R, 64, 90, interp, want_scores=False)
        with x360.Extractor(256, 512, 64, 90, R, mode) as ex:
            got, _ = ex.extract(frame[None])
        assert_same(got, want, f"active cameras {mode}")
    R30 = x360.GeometryProcessor.get_rotation_matrix(33.0, -20.0, 30.0)[None]   # roll honoured (geometry.py:112-116)
    want, _, _ = oracle.extract(frame[None], R30, 72, 100, 0, want_scores=False)
    with x360.Extractor(256, 512, 72, 100, R30, "linear") as ex:
        assert_same(ex.extract(frame[None])[0], want, "roll 30")


def test_chunked_host_pipeline_matches_single_shot(x360, oracle):
    """x360_extract_host splits the batch into max_batch chunks on side streams; ragged last chunk."""
    H, W, d = 128, 256, 64
    frames = np.stack([np.roll(noise_frame(i % 4, H, W), 37 * i, axis=1) for i in range(7)])
    R = rotations(x360, x360.GeometryProcessor.generate_views(6, 0, "cube"))
    want, ws, ws2 = oracle.extract(frames, R, d, 90, 0)
    with x360.Extractor(H, W, d, 90, R, "linear", max_batch=3) as ex:
        pinned = x360.PinnedBuffer(frames.shape)
        pinned.array[...] = frames
        got, s, s2 = ex.extract_sums(pinned.array)
        got2, s_b, s2_b = ex.extract_sums(pinned.array)           # plan reuse
    assert_same(got, want, "chunked pipeline")
    assert np.array_equal(s, ws) and np.array_equal(s2, ws2)
    assert np.array_equal(got, got2) and np.array_equal(s, s_b) and np.array_equal(s2, s2_b)
    with x360.Extractor(H, W, d, 90, R, "linear") as ex:
        out, sc = ex.extract(frames[:0], want_scores=True)         # empty batch
        assert out.shape == (0, 6, d, d, 3) and sc.shape == (0, 6)


def test_device_resident_path(x360, oracle):
    """torch tensors on the device, kernel enqueued on the caller's stream (what bench.py times)."""
    import torch
    H, W, d = 192, 384, 96
    frames = np.stack([smooth_frame(i, H, W) for i in range(4)])
    R = rotations(x360, x360.GeometryProcessor.generate_views(8, -20, "ring"))
    want, ws, ws2 = oracle.extract(frames, R, d, 90, 0)
    dev = torch.device("cuda", 0)
    d_frames = torch.from_numpy(frames).to(dev)
    with x360.Extractor(H, W, d, 90, R, "linear") as ex:
        d_out = torch.empty(ex.views_shape(4), dtype=torch.uint8, device=dev)
        d_s = torch.full((4, 8), -1, dtype=torch.int64, device=dev)
        d_s2 = torch.full((4, 8), -1, dtype=torch.int64, device=dev)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())             # uploads / fills above ran on the current stream
        with torch.cuda.stream(side):
            before = x360.launch_count()
            ex.extract_device(d_frames, d_out, d_s, d_s2, stream=side)
            assert x360.launch_count() in (before + 1, before + 2)          # one kernel; two where the plan defers pole tiles to a second launch
        side.synchronize()
        assert_same(d_out.cpu().numpy(), want, "device path")
        assert np.array_equal(d_s.cpu().numpy(), ws) and np.array_equal(d_s2.cpu().numpy(), ws2)
        ex.extract_device(d_frames, d_out, None, None, stream=torch.cuda.current_stream())
        torch.cuda.synchronize()
        assert_same(d_out.cpu().numpy(), want, "device path, no scores")
        with pytest.raises(ValueError, match="contiguous"):
            ex.extract_device(d_frames.transpose(1, 2).contiguous().transpose(1, 2), d_out)
        with pytest.raises(ValueError, match="shape"):
            ex.extract_device(d_frames[:, :100], d_out)


def test_explicit_map_remap_and_blur_score(x360, oracle):
    b = load_json("blur.json")
    edge = np.zeros((100, 100), np.uint8)
    edge[:, 50:] = 255
    assert x360.ImageUtils.calculate_blur_score(edge) == unhex(b["edge_100"]) == 1300.5      # test_core.py:84-93
    assert x360.ImageUtils.calculate_blur_score(np.ones((100, 100), np.uint8) * 128) == 0.0  # test_core.py:95-103
    assert x360.ImageUtils.calculate_blur_score(None) == 0.0                                 # test_core.py:105-108
    rng = np.random.default_rng(99)
    for name in ("color_100", "color_37x53", "gray_64x17", "color_2x2", "color_1x5", "color_5x1"):
        img = rng.integers(0, 256, b[name]["shape"], dtype=np.uint8)
        score = x360.ImageUtils.calculate_blur_score(img)
        assert isinstance(score, float) and score >= 0                                        # test_core.py:110-117
        assert score == pytest.approx(unhex(b[name]["score"]), rel=SCORE_RTOL, abs=1e-9)
    # cv2.remap drop-in with caller-held maps (processor.py:334)
    src = noise_frame(21, 120, 240)
    R = x360.GeometryProcessor.get_rotation_matrix(200.0, -35.0, 0.0)
    mx, my = x360.GeometryProcessor.create_rectilinear_map(120, 240, 70, 70, 95, 200.0, -35.0, 0.0)
    omx, omy = oracle.make_map(120, 240, 70, 70, 95, R)
    assert mx.shape == (70, 70) and mx.dtype == np.float32                                   # test_core.py:67-78
    assert np.array_equal(mx, omx) and np.array_equal(my, omy)
    for mode, interp in (("linear", 0), ("lanczos", 1)):
        assert_same(x360.remap(src, mx, my, mode), oracle.remap(src, omx, omy, interp), f"remap {mode}")
    # arbitrary (non-geometric) maps incl. out-of-range coordinates: BORDER_WRAP on both axes
    rng = np.random.default_rng(4)
    amx = rng.uniform(-300, 600, (40, 56)).astype(np.float32)
    amy = rng.uniform(-200, 400, (40, 56)).astype(np.float32)
    for mode, interp in (("linear", 0), ("lanczos", 1)):
        assert_same(x360.remap(src, amx, amy, mode), oracle.remap(src, amx, amy, interp), f"wild maps {mode}")


def test_session_mirrors_processor_loop(x360, oracle):
    """ExtractionSession = processor.py:245-265 + :327-376 for a frame batch."""
    settings = {"resolution": 64, "fov": 90, "camera_count": 6, "pitch_offset": -20, "layout_mode": "adaptive",
                "interpolation_mode": "linear", "blur_filter_enabled": True, "smart_blur_enabled": True,
                "blur_threshold": 500.0, "active_cameras": [0, 2, 4]}
    frames = np.stack([smooth_frame(1, 128, 256), noise_frame(2, 128, 256), smooth_frame(3, 128, 256)])
    with x360.ExtractionSession(settings, 128, 256) as sess:
        assert sess.names == ["View_0", "View_2", "View_4"]          # 'adaptive' => ring (processor.py:170-171)
        kept = sess.process(frames)
    views = x360.GeometryProcessor.generate_views(6, -20, "ring")
    _, R = x360.rotation_stack(views, [0, 2, 4])
    want, s, s2 = oracle.extract(frames, R, 64, 90, 0)
    scores = [[oracle.score_from_sums(s[f, v], s2[f, v], 64 * 64) for v in range(3)] for f in range(3)]
    ref_keep = x360.BlurFilter(True, True, 500.0).replay(scores)
    expect = [(f, ["View_0", "View_2", "View_4"][v]) for f in range(3) for v in range(3) if ref_keep[f][v]]
    assert [(k.frame_index, k.camera) for k in kept] == expect
    assert 0 < len(kept) < 9                                         # smooth frames rejected, noise kept
    for k in kept:
        v = ["View_0", "View_2", "View_4"].index(k.camera)
        assert_same(k.image, want[k.frame_index, v], "kept view")
        assert k.image.flags.c_contiguous and k.image.flags.writeable


@pytest.mark.parametrize("name", ["smart_ring4", "plain_threshold_cube", "smart_active_cameras", "filter_off_fibonacci",
                                  "smart_sharpened_lanczos", "legacy_adaptive_layout"])
def test_session_reproduces_the_reference_worker(x360, name):
    """ExtractionSession.process on the clip the REAL ProcessingWorker.process_video read (tests/golden/make_worker_golden.py):
    same kept (frame, camera) list, same scores (1e-12), and every kept view byte-identical to the PNG the worker saved —
    remap, blur score, accept / reject machine, sharpening of the accepted views, in the reference's order."""
    import json
    from test_host_api import worker_expectations
    with open(os.path.join(GOLDEN, "worker_cases.json")) as f:
        case = json.load(f)["scenarios"][name]
    frames = np.load(os.path.join(GOLDEN, "worker_clip.npz"))["frames"]
    names, table, kept = worker_expectations(x360, case)
    with x360.ExtractionSession(dict(case["settings"]), frames.shape[1], frames.shape[2], max_batch=3) as sess:
        assert sess.names == names
        got = sess.process(frames)
    assert [(k.frame_index, k.camera) for k in got] == kept
    for k in got:
        fn = f"clip_frame{k.frame_index:06d}_{k.camera}.png"
        assert sha(k.image) == case["kept"][fn], fn
        if table is not None:
            assert k.score == pytest.approx(table[k.frame_index, names.index(k.camera)], rel=SCORE_RTOL, abs=1e-9)


def test_filtered_download_keeps_state_across_calls(x360):
    """ExtractionSession.process downloads only the kept views (x360_extract_host_filtered) and carries the machine's state
    across calls exactly as one call over the whole clip does; the disabled filter keeps everything without scoring."""
    import json
    with open(os.path.join(GOLDEN, "worker_cases.json")) as f:
        case = json.load(f)["scenarios"]["smart_ring4"]
    frames = np.load(os.path.join(GOLDEN, "worker_clip.npz"))["frames"]
    with x360.ExtractionSession(dict(case["settings"]), frames.shape[1], frames.shape[2], max_batch=2) as sess:
        whole = sess.process(frames)
        skipped, forced = sess.blur.skipped, sess.blur.forced
    with x360.ExtractionSession(dict(case["settings"]), frames.shape[1], frames.shape[2], max_batch=3) as sess:
        parts = []
        for lo in (0, 5, 6, 11):
            hi = {0: 5, 5: 6, 6: 11, 11: len(frames)}[lo]
            parts += [(lo + k.frame_index, k.camera, sha(k.image)) for k in sess.process(frames[lo:hi])]
        assert (sess.blur.skipped, sess.blur.forced) == (skipped, forced) and forced >= 1
    assert parts == [(k.frame_index, k.camera, sha(k.image)) for k in whole]
    off = dict(case["settings"], blur_filter_enabled=False)
    with x360.ExtractionSession(off, frames.shape[1], frames.shape[2]) as sess:
        everything = sess.process(frames[:3])
    assert len(everything) == 3 * 4 and all(k.score is None for k in everything)


def test_blur_analyzer_mirrors_reference(x360, oracle):
    """BlurAnalyzer.analyze_frame = src/core/analyzer.py:44-76: ring layout whatever layout_mode says,
    bilinear, resolution default 1024; scores from the fused kernel == the oracle's to 1e-12."""
    frame = smooth_frame(5, 256, 512)
    settings = {"resolution": 96, "fov": 100, "camera_count": 5, "pitch_offset": -20, "layout_mode": "cube",
                "interpolation_mode": "lanczos"}
    res = x360.BlurAnalyzer.analyze_frame(frame, settings)
    views = x360.GeometryProcessor.generate_views(5, pitch_offset=-20)           # ring
    _, R = x360.rotation_stack(views)
    _, s, s2 = oracle.extract(frame[None], R, 96, 100, 0)
    want = [oracle.score_from_sums(s[0, v], s2[0, v], 96 * 96) for v in range(5)]
    assert [n for n, _ in res["details"]] == [v[0] for v in views]
    np.testing.assert_allclose([sc for _, sc in res["details"]], want, rtol=SCORE_RTOL)
    assert res["average"] == pytest.approx(sum(want) / 5, rel=SCORE_RTOL)
    assert res["min"] == pytest.approx(min(want), rel=SCORE_RTOL) and res["max"] == pytest.approx(max(want), rel=SCORE_RTOL)
    assert all(isinstance(sc, float) for _, sc in res["details"])


LZ_COL_CASES = [
    # (H, W, d, fov, n, F): Lanczos4 on yaw-only plans (a thread's pixels read the same source columns; weights computed in registers)
    (512, 1024, 256, 90, 8, 3),        # 1.27 source px per output px
    (720, 1440, 200, 90, 4, 2),        # 2.3: windows 2.3 rows apart, partial tiles (200 = 6 x 32 + 8)
    (300, 600, 250, 60, 3, 2),         # magnification: consecutive pixels share all eight rows; frame pitch not a multiple of 16 B
    (1440, 2880, 333, 100, 5, 1),      # odd size, seam view at yaw 144 / 216
    (480, 960, 64, 120, 2, 2),         # 4.3 source px per output px: beyond the staging capacity on most tiles
]


@pytest.mark.parametrize("scores", [False, True], ids=["noscores", "fused-score"])
@pytest.mark.parametrize("cfg", LZ_COL_CASES)
def test_lanczos_ring_plans(x360, oracle, cfg, scores):
    H, W, d, fov, n, F = cfg
    frames = np.stack([noise_frame(700 + i, H, W) for i in range(F)])
    R = rotations(x360, x360.GeometryProcessor.generate_views(n, 0, "ring"))
    assert x360.column_only_map_x(R)
    want, ws, ws2 = oracle.extract(frames, R, d, fov, 1, want_scores=scores)
    with x360.Extractor(H, W, d, fov, R, "lanczos", max_batch=F, flags=x360.FLAG_FUSED_SCORE if scores else 0) as ex:
        if scores:
            got, s, s2 = ex.extract_sums(frames)
            assert np.array_equal(s, ws) and np.array_equal(s2, ws2), "integer Laplacian sums"
        else:
            got, _ = ex.extract(frames, want_scores=False)
    assert_same(got, want, f"{cfg} lanczos ring", LANCZOS_TOL_LSB)


# ---- plan variants and robustness of the host side ----------------------------------------------
POLE_CASES = [
    # (H, W, d, fov, layout, n, pitch, F): sources wide enough for the seam window (3 W >= 2048) and pole faces under small views
    (512, 1024, 192, 90, "cube", 6, 0, 3),
    (640, 1280, 160, 90, "cube", 6, -30, 2),
    (768, 1536, 224, 100, "fibonacci", 9, -20, 2),
    (720, 1440, 200, 90, "ring", 5, 70, 2),             # inclined ring: every view sees the pole region
]


@pytest.mark.parametrize("scores", [False, True], ids=["noscores", "scores"])
@pytest.mark.parametrize("cfg", POLE_CASES)
def test_wide_seam_and_second_launch_against_oracle(x360, oracle, cfg, scores):
    """Tiles near the poles (wide boxes, the second launch with larger buffers) and across the +-pi seam (seam window) are
    staged instead of gathered from global memory: same bytes as the oracle and as the plan without them (FLAG_NO_WIDE)."""
    H, W, d, fov, layout, n, pitch, F = cfg
    frames = np.stack([noise_frame(900 + i, H, W) for i in range(F)])
    R = rotations(x360, x360.GeometryProcessor.generate_views(n, pitch_offset=pitch, layout_mode=layout))
    want, ws, ws2 = oracle.extract(frames, R, d, fov, 0, want_scores=scores)
    results = {}
    for name, flags in (("auto", 0), ("no_wide", x360.FLAG_NO_WIDE), ("raw_fused", x360.FLAG_FUSED_SCORE),
                        ("planes_fused", x360.FLAG_FUSED_SCORE | x360.FLAG_PLANES), ("split", x360.FLAG_SPLIT_SCORE)):
        if not scores and name in ("raw_fused", "planes_fused", "split"):
            continue
        with x360.Extractor(H, W, d, fov, R, "linear", max_batch=F, flags=flags) as ex:
            if scores:
                got, s, s2 = ex.extract_sums(frames)
                assert np.array_equal(s, ws) and np.array_equal(s2, ws2), f"{cfg} {name}: integer Laplacian sums"
            else:
                got, _ = ex.extract(frames, want_scores=False)
            results[name] = ex.tile_modes()
        assert_same(got, want, f"{cfg} {name}")
    auto, nw = results["auto"], results["no_wide"]
    assert sum(auto.values()) == sum(nw.values())
    assert nw["staged_wide"] == 0 and nw["staged_seam"] == 0
    direct = lambda m: m["direct_aligned"] + m["direct_bytewise"]
    assert direct(auto) < direct(nw), (auto, nw)                        # wide / seam / second launch took tiles off the direct path
    if not scores:
        assert auto["staged_seam"] > 0, auto


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", [(512, 1024, 192, 90, "cube", 6, 0, 3), (640, 1280, 160, 90, "cube", 6, -30, 2), (720, 1440, 224, 90, "ring", 5, 70, 2),
                                 (1024, 2048, 512, 90, "cube", 6, 0, 2), (512, 1024, 96, 120, "cube", 6, 0, 1)])
def test_transposed_tiles_against_oracle(x360, oracle, cfg):
    """Square views of whole 32 x 32 blocks with 16-row tiles: the blocks of a general view where the source row changes faster
    along x than along y are gathered transposed (lanes down the rows).  Same bytes as the oracle, as the 32-row plan (which has
    no such tiles), with TMA stores and without (odd output pointer)."""
    H, W, d, fov, layout, n, pitch, F = cfg
    frames = np.stack([noise_frame(1300 + i, H, W) for i in range(F)])
    R = rotations(x360, x360.GeometryProcessor.generate_views(n, pitch_offset=pitch, layout_mode=layout))
    want, _, _ = oracle.extract(frames, R, d, fov, 0, want_scores=False)
    with x360.Extractor(H, W, d, fov, R, "linear", max_batch=F, flags=x360.FLAG_TILE_H16) as ex:
        assert ex.transposed_blocks() > 0, cfg
        assert_same(ex.extract(frames, want_scores=False)[0], want, f"{cfg} 16-row tiles, transposed blocks")
        got, s, s2 = ex.extract_sums(frames)                       # the score variants never transpose: same views all the same
        assert_same(got, want, f"{cfg} with the split score")
        import torch
        dev_frames = torch.from_numpy(frames).cuda()
        spare = torch.empty(F * len(R) * d * d * 3 + 1, dtype=torch.uint8, device="cuda")
        views = spare[1:].view(F, len(R), d, d, 3)                  # misaligned output: byte stores instead of TMA stores
        ex.extract_device(dev_frames, views)
        torch.cuda.synchronize()
        assert_same(views.cpu().numpy(), want, f"{cfg} transposed blocks, byte stores")
    with x360.Extractor(H, W, d, fov, R, "linear", max_batch=F, flags=x360.FLAG_TILE_H32) as ex:
        assert ex.transposed_blocks() == 0
        assert_same(ex.extract(frames, want_scores=False)[0], want, f"{cfg} 32-row tiles")


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", [(1024, 2048, 512, 90, "cube", 6, 0, 1), (960, 1920, 448, 100, "fibonacci", 9, -20, 2)])
def test_lanczos_second_launch_against_oracle(x360, oracle, cfg):
    """Lanczos4 plans defer the (tile, view) pairs that fit no staging box to a second launch with larger buffers (from one wave
    of them on): same bytes as the oracle and as the plan without it (FLAG_NO_WIDE), fewer pairs on the direct gathers."""
    H, W, d, fov, layout, n, pitch, F = cfg
    frames = np.stack([noise_frame(1500 + i, H, W) for i in range(F)])
    R = rotations(x360, x360.GeometryProcessor.generate_views(n, pitch_offset=pitch, layout_mode=layout))
    want, ws, ws2 = oracle.extract(frames, R, d, fov, 1, want_scores=True)
    modes = {}
    for name, flags in (("auto", 0), ("no_wide", x360.FLAG_NO_WIDE)):
        with x360.Extractor(H, W, d, fov, R, "lanczos", max_batch=F, flags=flags) as ex:
            got, _ = ex.extract(frames, want_scores=False)
            modes[name] = ex.tile_modes()
            assert_same(got, want, f"{cfg} {name}", LANCZOS_TOL_LSB)
            got, s, s2 = ex.extract_sums(frames)
            assert_same(got, want, f"{cfg} {name} with scores", LANCZOS_TOL_LSB)
            assert np.array_equal(s, ws) and np.array_equal(s2, ws2), f"{cfg} {name}: integer Laplacian sums"
    direct = lambda m: m["direct_aligned"] + m["direct_bytewise"]
    assert direct(modes["auto"]) < direct(modes["no_wide"]), modes


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["linear", "lanczos"])
def test_host_path_uploads_only_the_rows_the_views_read(x360, oracle, mode):
    """A ring at pitch 0 reads a band of source rows: the host entries upload that band only (x360_plan_upload_rows).  The band
    is sufficient — same views as the oracle — and nothing outside it is read: zeroing the other rows of the frames given to the
    library changes nothing.  Plans that see a pole, and FLAG_FULL_UPLOAD, take whole frames."""
    H, W, d, fov, F = 720, 1440, 200, 90, 5
    frames = np.stack([noise_frame(1700 + i, H, W) for i in range(F)])
    R = rotations(x360, x360.GeometryProcessor.generate_views(6, 0, "ring"))
    interp = 1 if mode == "lanczos" else 0
    want, ws, ws2 = oracle.extract(frames, R, d, fov, interp, want_scores=True)
    with x360.Extractor(H, W, d, fov, R, mode, max_batch=2) as ex:
        r0, n = ex.upload_rows()
        assert 0 < r0 and r0 + n < H and n < 0.75 * H, (r0, n)
        poisoned = frames.copy()
        poisoned[:, :r0] = 0
        poisoned[:, r0 + n:] = 0
        for src in (frames, poisoned):
            got, s, s2 = ex.extract_sums(src)
            assert_same(got, want, f"{mode} band [{r0}, {r0 + n})", LANCZOS_TOL_LSB if interp else 0)
            assert np.array_equal(s, ws) and np.array_equal(s2, ws2)
        files, _ = ex.extract_jpeg(poisoned)
        assert cv2_decode_equal(files[F - 1][len(R) - 1], want[F - 1, len(R) - 1])
    with x360.Extractor(H, W, d, fov, R, mode, max_batch=2, flags=x360.FLAG_FULL_UPLOAD) as ex:
        assert ex.upload_rows() == (0, H)
        assert_same(ex.extract(frames, want_scores=False)[0], want, f"{mode} full upload", LANCZOS_TOL_LSB if interp else 0)
    Rc = rotations(x360, x360.GeometryProcessor.generate_views(6, 0, "cube"))
    with x360.Extractor(H, W, d, fov, Rc, mode, max_batch=2) as ex:
        assert ex.upload_rows() == (0, H)                               # the pole faces read every row


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["linear", "lanczos"])
def test_single_frame_jobs_are_pipelined_by_view_range(x360, oracle, mode):
    """Plans whose views share no map (fibonacci) extract a single-frame chunk in up to eight launches over consecutive view ranges and
    download each range while the next is computed (x360_extract_host): same bytes, whatever the chunking."""
    H, W, d, fov = 768, 1536, 224, 100
    interp = 1 if mode == "lanczos" else 0
    for n_views in (9, 4, 3):                                        # 3 views: fewer than the four ranges, one launch
        R = rotations(x360, x360.GeometryProcessor.generate_views(n_views, pitch_offset=-20, layout_mode="fibonacci"))
        frames = np.stack([noise_frame(1900 + i, H, W) for i in range(3)])
        want, _, _ = oracle.extract(frames, R, d, fov, interp, want_scores=False)
        with x360.Extractor(H, W, d, fov, R, mode, max_batch=1) as ex:             # chunks of one frame
            assert_same(ex.extract(frames, want_scores=False)[0], want, f"{mode} {n_views} views, three one-frame chunks", LANCZOS_TOL_LSB if interp else 0)
            assert_same(ex.extract(frames[1:2], want_scores=False)[0], want[1:2], f"{mode} {n_views} views, one frame", LANCZOS_TOL_LSB if interp else 0)
        with x360.Extractor(H, W, d, fov, R, mode, max_batch=4) as ex:             # a 3-frame chunk: one launch for all views
            assert_same(ex.extract(frames, want_scores=False)[0], want, f"{mode} {n_views} views, one chunk", LANCZOS_TOL_LSB if interp else 0)


def cv2_decode_equal(jpeg_bytes, view):
    """The file decodes (cv2) to what cv2's own encode of the view decodes to."""
    import cv2
    ok, ref = cv2.imencode(".jpg", view, [cv2.IMWRITE_JPEG_QUALITY, 95])
    return ok and bytes(ref.tobytes()) == bytes(jpeg_bytes)


def test_two_live_plans_do_not_break_each_other(x360, oracle):
    """The dynamic shared-memory limit belongs to the kernel function: a second, smaller plan must not lower it under a live
    larger one (BlurAnalyzer / create_rectilinear_map build temporary plans while a session is alive)."""
    H, W = 512, 1024
    frames = np.stack([noise_frame(40 + i, H, W) for i in range(2)])
    R = rotations(x360, x360.GeometryProcessor.generate_views(6, 0, "ring"))
    big = x360.Extractor(H, W, 320, 110, R, "linear", max_batch=2)
    want_big, _, _ = oracle.extract(frames, R, 320, 110, 0, want_scores=False)
    assert_same(big.extract(frames)[0], want_big, "first plan, first call")
    with x360.Extractor(H, W, 64, 60, R[:2], "linear") as small:                     # same kernel variant, much smaller footprint
        want_small, _, _ = oracle.extract(frames, R[:2], 64, 60, 0, want_scores=False)
        assert_same(small.extract(frames)[0], want_small, "second plan")
        assert_same(big.extract(frames)[0], want_big, "first plan while the second is alive")
    assert_same(big.extract(frames)[0], want_big, "first plan after the second was destroyed")
    big.close()


def test_launches_on_two_streams_do_not_share_scratch(x360, oracle):
    """x360_extract_device is asynchronous on any stream: every launch gets its own work-item counter."""
    import torch
    H, W, d = 384, 768, 256
    R = rotations(x360, x360.GeometryProcessor.generate_views(8, 0, "ring"))
    fa = np.stack([noise_frame(60 + i, H, W) for i in range(6)])
    fb = np.stack([noise_frame(70 + i, H, W) for i in range(6)])
    want_a, _, _ = oracle.extract(fa, R, d, 90, 0, want_scores=False)
    want_b, _, _ = oracle.extract(fb, R, d, 90, 0, want_scores=False)
    with x360.Extractor(H, W, d, 90, R, "linear") as ex:
        da, db = torch.from_numpy(fa).cuda(), torch.from_numpy(fb).cuda()
        oa = torch.empty(ex.views_shape(6), dtype=torch.uint8, device="cuda")
        ob = torch.empty_like(oa)
        sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
        torch.cuda.synchronize()
        for _ in range(5):                                             # back to back on two streams: the kernels overlap
            ex.extract_device(da, oa, stream=sa)
            ex.extract_device(db, ob, stream=sb)
        torch.cuda.synchronize()
        assert_same(oa.cpu().numpy(), want_a, "stream A")
        assert_same(ob.cpu().numpy(), want_b, "stream B")


def test_extract_sums_validates_like_extract(x360):
    H, W = 128, 256
    R = rotations(x360, x360.GeometryProcessor.generate_views(2, 0, "ring"))
    with x360.Extractor(H, W, 64, 90, R, "linear") as ex:
        for fn in (ex.extract, ex.extract_sums):
            with pytest.raises(ValueError):
                fn(np.zeros((1, H, W + 2, 3), np.uint8))                               # wrong frame size
            with pytest.raises(TypeError):
                fn(np.zeros((1, H, W, 3), np.float32))                                 # silently casting would hide a caller bug
            with pytest.raises(ValueError):
                fn(np.zeros((1, H, W, 3), np.uint8), out=np.zeros((1, 2, 64, 64, 4), np.uint8))
            with pytest.raises(ValueError):
                fn(np.zeros((1, H, W, 3), np.uint8), out=np.zeros((1, 2, 64, 64, 3), np.uint16))
    with pytest.raises(ValueError):
        x360.ExtractionSession({"is_360": False}, H, W)                                # flat media are not reprojected


def test_fuzz_seed_against_oracle(x360, oracle):
    """One reduced seed of scripts/fuzz_parity.py inside the GPU suite (random geometries, flags, chunking): all bit-exact."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
    import fuzz_parity
    rep = fuzz_parity.run(n_cases=60, seed=4242, verbose=False)
    assert rep["cases"] == 60 and rep["all_exact"], rep["failures"][:3]


# ---- device JPEG encode (the save step: processor.py:121-123, file_manager.py:21-32) -------------------------------------
JPEG_GPU_CASES = [(16, 16, 95, 1), (48, 64, 95, 3), (100, 100, 95, 2), (17, 33, 75, 2), (233, 77, 50, 1), (256, 256, 95, 5), (8, 8, 100, 1),
                  (1, 1, 95, 1), (24, 40, 30, 2), (640, 640, 95, 2)]


@pytest.mark.parametrize("cfg", JPEG_GPU_CASES)
def test_jpeg_files_are_byte_identical_to_cv2(x360, oracle, cfg):
    """JpegEncoder.encode == cv2.imencode(.jpg, quality) == the oracle, byte for byte; the quantised coefficients too."""
    cv2 = pytest.importorskip("cv2")
    h, w, q, n = cfg
    rng = np.random.default_rng(h * 1000 + w)
    imgs = rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)
    if h >= 64:
        lo = rng.integers(0, 256, (h // 8 + 1, w // 8 + 1, 3), dtype=np.uint8)
        imgs[0] = cv2.resize(lo, (w, h), interpolation=cv2.INTER_CUBIC)                # camera-like content
    imgs[-1, : h // 2] = 255                                                            # a saturated band: long runs of 0xFF / zeros
    with x360.JpegEncoder(h, w, q, max_images=2) as enc:                                # n > max_images: chunked
        coef = enc.coefficients(imgs[:2])
        files = enc.encode(imgs)
    assert np.array_equal(coef, np.stack([oracle.jpeg_coefficients(im, q) for im in imgs[:2]]))
    for i, im in enumerate(imgs):
        want = cv2.imencode(".jpg", im, [cv2.IMWRITE_JPEG_QUALITY, q])[1].tobytes()
        assert files[i] == want, f"{cfg} image {i}: {len(files[i])} vs {len(want)} bytes"
        assert files[i] == oracle.jpeg_encode(im, q)


def test_jpeg_device_entry_and_capacity_check(x360):
    cv2 = pytest.importorskip("cv2")
    import torch
    rng = np.random.default_rng(8)
    imgs = rng.integers(0, 256, (3, 96, 128, 3), dtype=np.uint8)
    with x360.JpegEncoder(96, 128, 95, max_images=4) as enc:
        d = torch.from_numpy(imgs).cuda()
        out = torch.empty(enc.max_bytes(3), dtype=torch.uint8, device="cuda")
        off = torch.zeros(4, dtype=torch.int64, device="cuda")
        st = torch.cuda.Stream()
        st.wait_stream(torch.cuda.current_stream())
        enc.encode_device(d, out, off, stream=st)
        enc.status(st)
        o = off.cpu().numpy()
        blob = out.cpu().numpy()
        for i in range(3):
            assert blob[o[i]:o[i + 1]].tobytes() == cv2.imencode(".jpg", imgs[i], [cv2.IMWRITE_JPEG_QUALITY, 95])[1].tobytes()
        small = torch.empty(1000, dtype=torch.uint8, device="cuda")                    # three noise files need ~45 KB
        enc.encode_device(d, small, off, stream=st)
        with pytest.raises(x360.X360Error):
            enc.status(st)
        enc.encode_device(d, out, off, stream=st)                                       # the encoder is usable again
        enc.status(st)


@pytest.mark.parametrize("sharpen", [False, True], ids=["plain", "sharpened"])
def test_extract_jpeg_equals_cv2_encode_of_the_views(x360, sharpen):
    """Extractor.extract_jpeg / ExtractionSession.process_encoded: the pipelined host path with the encode on the device returns,
    per (frame, view), exactly the bytes cv2.imwrite would put in the file for the view that extract() returns."""
    cv2 = pytest.importorskip("cv2")
    H, W = 256, 512
    frames = np.stack([smooth_frame(30 + i, H, W) for i in range(5)])
    frames[3] = noise_frame(33, H, W)
    settings = {"resolution": 160, "fov": 90, "camera_count": 6, "layout_mode": "cube", "interpolation_mode": "linear", "quality": 92,
                "blur_filter_enabled": True, "smart_blur_enabled": False, "blur_threshold": 50.0,
                "sharpening_enabled": sharpen, "sharpening_strength": 0.6}
    with x360.ExtractionSession(settings, H, W, max_batch=2) as sess:
        kept = sess.process(frames)
    with x360.ExtractionSession(settings, H, W, max_batch=2) as sess:
        enc = sess.process_encoded(frames)
        files, scores = sess.extractor.extract_jpeg(frames[:2], quality=92, want_scores=True)
    assert [(k.frame_index, k.camera) for k in kept] == [(f, c) for f, c, _, _ in enc] and len(kept) > 0
    for k, (f, c, data, score) in zip(kept, enc):
        assert data == cv2.imencode(".jpg", k.image, [cv2.IMWRITE_JPEG_QUALITY, 92])[1].tobytes(), (f, c)
        assert score == pytest.approx(k.score, rel=SCORE_RTOL, abs=1e-9)
    assert len(files) == 2 and len(files[0]) == 6 and scores.shape == (2, 6)


# ---- BASELINE.json configurations at full size ------------------------------------------------
FULL = [
    # name, (H, W), layout, n, pitch, active, d, fov, mode, F, views to check against the oracle
    ("cfg1", (1920, 3840), "cube", 6, 0, None, 1024, 90, "linear", 1, [0, 1, 2, 3, 4, 5]),
    ("cfg2", (2880, 5760), "ring", 8, 0, None, 1600, 90, "linear", 2, [0, 1, 2, 3, 4, 5, 6, 7]),   # all views: one shared map
    ("cfg3", (3840, 7680), "fibonacci", 20, -20, None, 1920, 90, "lanczos", 1, [0, 7, 11, 19]),
    ("cfg4", (3840, 7680), "cube", 6, 0, None, 1920, 90, "linear", 2, [0, 1, 2, 3, 4, 5]),
    ("cfg5", (5760, 11520), "ring", 12, 0, [0, 3, 6, 9], 2048, 90, "lanczos", 1, [0, 1, 2, 3]),
]


@pytest.mark.parametrize("cfg", FULL, ids=[c[0] for c in FULL])
def test_full_size_configs(x360, oracle, cfg):
    name, (H, W), layout, n, pitch, active, d, fov, mode, F, check = cfg
    interp = 1 if mode == "lanczos" else 0
    base = noise_frame(0, H, W)
    frames = np.stack([np.roll(base, 37 * i, axis=1) for i in range(F)])       # SURVEY.md 8(d) clip recipe
    views = x360.GeometryProcessor.generate_views(n, pitch_offset=pitch, layout_mode=layout)
    names, R = x360.rotation_stack(views, active)
    with x360.Extractor(H, W, d, fov, R, mode, max_batch=1) as ex:
        got, s, s2 = ex.extract_sums(frames)
        modes = ex.tile_modes()
        again, s_b, s2_b = ex.extract_sums(frames)
    if not interp:                                      # most tiles take the staged path; poles / seam fall back
        assert modes["staged"] + modes["staged_wide"] + modes["staged_seam"] > 0.7 * sum(modes.values()), modes
    # size-independent properties: deterministic, every view differs, sums consistent with the pixels
    assert np.array_equal(got, again) and np.array_equal(s, s_b) and np.array_equal(s2, s2_b)
    assert len({sha(got[0, v]) for v in range(len(names))}) == len(names)
    f = F - 1
    for v in check:
        want, ws, ws2 = oracle.extract(frames[f:f + 1], R[v:v + 1], d, fov, interp)
        assert_same(got[f, v], want[0, 0], f"{name} frame {f} view {names[v]}", LANCZOS_TOL_LSB if interp else 0)
        assert s[f, v] == ws[0, 0] and s2[f, v] == ws2[0, 0]
        assert oracle.blur_sums(got[f, v]) == (s[f, v], s2[f, v])


# ---- unsharp mask (processor.py:379-381) ------------------------------------------------------------------

@pytest.mark.gpu
def test_unsharp_matches_golden_and_oracle(x360, oracle):
    """Committed cv2 outputs (tests/golden/make_golden.py) and the C oracle, bit-exact; the golden cases are small,
    so they run on the any-size kernel, the seeded larger ones on the TMA-tiled kernel (both are checked)."""
    u = load_json("unsharp.json")
    z = np.load(os.path.join(GOLDEN, "unsharp_cases.npz"))
    for k, m in enumerate(u["cases"]):
        img, s = z[f"in{k}"], unhex(m["strength"])
        assert np.array_equal(x360.unsharp_mask(img, s), z[f"out{k}"]), m["name"]
        assert np.array_equal(x360.unsharp_mask(img, -1.0), z[f"blur{k}"]), m["name"] + " (blur alone)"
        assert np.array_equal(x360.unsharp_mask(img, 0.0), img)
    rng = np.random.default_rng(77)
    for (n, h, w), s in [((3, 96, 96), 0.5), ((2, 44, 112), 1.3), ((1, 130, 208), 0.3), ((2, 97, 128), 2.0), ((1, 45, 96), 0.05),
                         ((1, 64, 100), 0.5), ((2, 7, 96), 0.7)]:
        imgs = rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)
        imgs[0, : h // 2] //= 3                                    # a darker band: fewer saturated outputs
        got = x360.unsharp_mask(imgs, s)
        want = np.stack([oracle.unsharp(im, s) for im in imgs])
        assert np.array_equal(got, want), f"{(n, h, w)} strength {s}: {(got != want).sum()} bytes differ"


@pytest.mark.gpu
def test_unsharp_paths_agree_at_full_size(x360, oracle):
    """1600 x 1600 views (cfg2's size): tiled kernel == any-size kernel == oracle on one view; device entry == host entry."""
    import torch
    rng = np.random.default_rng(5)
    small = rng.integers(0, 256, (2, 200, 200, 3), dtype=np.uint8)
    imgs = np.ascontiguousarray(np.repeat(np.repeat(small, 8, axis=1), 8, axis=2))
    imgs[:, ::3, 1::2] ^= 0x55                                     # blocky texture + high-frequency detail
    tiled = x360.unsharp_mask(imgs, 0.5)
    assert np.array_equal(tiled[1], oracle.unsharp(imgs[1], 0.5))
    odd = np.ascontiguousarray(imgs[1:, :, :1599])                 # row pitch not a multiple of 16 B: the any-size kernel
    assert np.array_equal(x360.unsharp_mask(odd, 0.5)[0], oracle.unsharp(odd[0], 0.5))
    d_in = torch.from_numpy(imgs).cuda()
    d_out = torch.empty_like(d_in)
    x360.unsharp_mask_device(d_in, d_out, 0.5)
    torch.cuda.synchronize()
    assert np.array_equal(d_out.cpu().numpy(), tiled)
    with pytest.raises(x360.X360Error):
        x360.unsharp_mask_device(d_in, d_in, 0.5)                  # in place is refused


@pytest.mark.gpu
def test_session_sharpens_accepted_views_only(x360, oracle):
    """ExtractionSession.process with sharpening_enabled == the reference order: score on the unsharpened view,
    reject, then sharpen what was kept (processor.py:341-381)."""
    rng = np.random.default_rng(11)
    frames = rng.integers(0, 256, (2, 128, 256, 3), dtype=np.uint8)
    base = {"resolution": 96, "fov": 90, "camera_count": 4, "layout_mode": "ring", "blur_filter_enabled": True,
            "blur_threshold": 100.0}
    with x360.ExtractionSession(dict(base), 128, 256) as plain:
        ref = plain.process(frames)
    with x360.ExtractionSession(dict(base, sharpening_enabled=True, sharpening_strength=0.8), 128, 256) as sharp:
        got = sharp.process(frames)
    assert [(r.frame_index, r.camera, r.score) for r in got] == [(r.frame_index, r.camera, r.score) for r in ref]
    assert len(got) > 0
    for g, r in zip(got, ref):
        assert np.array_equal(g.image, oracle.unsharp(r.image, 0.8))


@pytest.mark.gpu
def test_extract_host_with_sharpening_is_extract_then_unsharp(x360, oracle):
    """Extractor(sharpen_strength=s): the pipelined host path (5 frames in chunks of 2) returns the unsharp-masked views,
    the sums stay those of the plain views; switching it off restores the plain output."""
    rng = np.random.default_rng(21)
    frames = rng.integers(0, 256, (5, 128, 256, 3), dtype=np.uint8)
    _, R = x360.rotation_stack(x360.GeometryProcessor.generate_views(6, layout_mode="cube"))
    with x360.Extractor(128, 256, 96, 90, R, "linear", max_batch=2) as ex:
        plain, s, s2 = ex.extract_sums(frames)
        ex.set_sharpen(0.6)
        sharp, t, t2 = ex.extract_sums(frames)
        ex.set_sharpen(None)
        again, _ = ex.extract(frames)
    assert np.array_equal(again, plain)
    assert np.array_equal(s, t) and np.array_equal(s2, t2)
    for f in range(5):
        for v in range(6):
            assert np.array_equal(sharp[f, v], oracle.unsharp(plain[f, v], 0.6)), (f, v)


# ---- device-resident hand-off to the masker (processor.py:392-420) ------------------------------------------------

@pytest.mark.gpu
@pytest.mark.parametrize("d,n,F", [(64, 6, 2), (50, 6, 1), (33, 36, 3)])     # 33 / 50: the any-size kernel; 36 x 3 = 108 slots: two launches
def test_masker_handoff_stays_on_the_device(x360, oracle, d, n, F):
    torch = pytest.importorskip("torch")
    from oracle import cv2_path
    H, W = 128, 256
    layout = "cube" if n == 6 else "ring"
    views = x360.GeometryProcessor.generate_views(n, 0, layout)
    names, R = x360.rotation_stack(views)
    frames = np.stack([noise_frame(700 + i, H, W) for i in range(F)])
    want_views, _, _ = oracle.extract(frames, R, d, 90, 0, want_scores=False)
    fr = torch.from_numpy(frames).cuda()
    out = torch.empty((F, len(names), d, d, 3), dtype=torch.uint8, device="cuda")
    with x360.Extractor(H, W, d, 90, R, "linear") as ex:
        ex.extract_device(fr, out, stream=torch.cuda.current_stream())
        for cams, rgb in ((None, True), ("Up,down" if n == 6 else "view_1, VIEW_7", True), (None, False)):
            slots, t = x360.masker_input(out, names, cams, to_rgb=rgb)
            torch.cuda.synchronize()
            flat = want_views.reshape(F * len(names), d, d, 3)
            assert slots == x360.masker_slots(F, names, cams)
            assert t.shape == (len(slots), 3, d, d) and t.dtype == torch.float32 and t.is_cuda
            want = cv2_path.masker_tensor(flat[slots], to_rgb=rgb) if slots else np.empty((0, 3, d, d), np.float32)
            assert np.array_equal(t.cpu().numpy(), want), f"hand-off differs (cams={cams}, rgb={rgb})"
    if n == 6:
        assert x360.masker_slots(F, names, "Up,down") == [f * 6 + v for f in range(F) for v in (4, 5)]
