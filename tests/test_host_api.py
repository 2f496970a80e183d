"""Host layer (no GPU): layouts, rotations, the C-ABI surface, the blur state machine.

The TestGeometryProcessor block reads like the reference's tests/test_core.py:18-65 on purpose:
same calls, same expectations, against the drop-in class."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, load_json, unhex


# ---- reference tests/test_core.py:18-65, transposed onto the drop-in -------------------------
def test_generate_views_ring_6(x360):
    views = x360.GeometryProcessor.generate_views(6, pitch_offset=0, layout_mode="ring")
    assert len(views) == 6
    for (_, yaw, _, _), expected in zip(views, [0.0, 60.0, 120.0, 180.0, 240.0, 300.0]):
        assert yaw == pytest.approx(expected, abs=0.05)


def test_generate_views_cube(x360):
    views = x360.GeometryProcessor.generate_views(12, layout_mode="cube")
    assert len(views) == 6
    names = [v[0] for v in views]
    for n in ("Front", "Back", "Up", "Down"):
        assert n in names


def test_generate_views_fibonacci(x360):
    for n in (4, 8, 16):
        assert len(x360.GeometryProcessor.generate_views(n, layout_mode="fibonacci")) == n


def test_generate_views_with_pitch_offset(x360):
    for _, _, pitch, _ in x360.GeometryProcessor.generate_views(4, pitch_offset=-20, layout_mode="ring"):
        assert pitch == -20


def test_rotation_matrix_identity(x360):
    R = x360.GeometryProcessor.get_rotation_matrix(0, 0, 0)
    assert R.shape == (3, 3)
    np.testing.assert_array_almost_equal(R, np.eye(3), decimal=5)


# ---- bit-level pins from the real reference ---------------------------------------------------
def test_layouts_bit_identical_to_reference(x360):
    g = load_json("geometry.json")
    for key, exp in g["views"].items():
        layout, n, pitch = key.split(",")
        got = x360.GeometryProcessor.generate_views(int(n), pitch_offset=int(pitch), layout_mode=layout)
        want = [(a, unhex(b), unhex(c), d) for a, b, c, d in exp]
        assert got == want, key


def test_rotations_bit_identical_to_reference(x360):
    g = load_json("geometry.json")
    for key, exp in g["rotations"].items():
        ypr = [unhex(s) for s in key.split(",")]
        R = x360.GeometryProcessor.get_rotation_matrix(*ypr)
        assert R.dtype == np.float64
        assert [float(v).hex() for v in R.reshape(-1)] == exp, key
    # the residues that decide the seam
    R = x360.GeometryProcessor.get_rotation_matrix(180, 0, 0)
    assert R[0, 2] == 1.2246467991473532e-16 and R[2, 0] == -1.2246467991473532e-16


def test_rotation_stack_active_cameras(x360):
    views = x360.GeometryProcessor.generate_views(12, 0, "ring")
    names, R = x360.rotation_stack(views, active_cameras=[0, 3, 6, 9])
    assert names == ["View_0", "View_3", "View_6", "View_9"] and R.shape == (4, 3, 3)
    assert np.array_equal(R[1], x360.GeometryProcessor.get_rotation_matrix(90.0, 0, 0))
    names, R = x360.rotation_stack(views, active_cameras=[])
    assert names == [] and R.shape == (0, 3, 3)


# ---- C-ABI surface ----------------------------------------------------------------------------
def test_library_exports_every_declared_symbol(x360):
    """Every function include/x360.h declares is exported by libx360.so and bound by _capi."""
    hdr = open(os.path.join(ROOT, "include", "x360.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(x360_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 16
    lib = ctypes.CDLL(x360._capi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in x360.h but not exported"
    assert declared == set(x360._capi.SYMBOLS), "ctypes binding and header disagree"
    assert x360.version().startswith("x360 ")
    assert x360.launch_count() >= 0


def test_params_struct_layout_matches_header(x360):
    P = x360._capi.Params
    assert ctypes.sizeof(P) == 64            # 5*i32 (+4 pad) + f64 + 2*i32 + ptr + 4*i32
    assert P.fov_deg.offset == 24 and P.R.offset == 40 and P.device.offset == 48


def test_host_side_tables_match_oracle(x360, oracle):
    tab = np.empty((1024, 8, 8), np.int16)
    assert x360._capi.lib().x360_lanczos4_table(tab.ctypes.data) == 0
    assert np.array_equal(tab, oracle.lanczos4_table())
    assert (tab.reshape(1024, -1).astype(np.int64).sum(1) == 32768).all()
    assert tab.min() == -5727 and tab.max() == 32767
    for w, fov in [(1024, 90), (1600, 90), (1920, 110), (2048, 60), (1023, 120), (512, 73.5)]:
        ref = (0.5 * w) / np.tan(0.5 * np.radians(fov))          # geometry.py:141
        assert x360.focal_length(w, fov) == ref == oracle.focal(w, fov)


def test_argument_errors_do_not_need_a_gpu(x360):
    lib = x360._capi.lib()
    assert lib.x360_plan_create(None, None) == -1
    assert b"null" in lib.x360_last_error()
    p = x360._capi.Params()
    p.struct_size = 12
    plan = ctypes.c_void_p()
    assert lib.x360_plan_create(ctypes.byref(p), ctypes.byref(plan)) == -1
    assert b"ABI mismatch" in lib.x360_last_error()
    with pytest.raises(x360.X360Error):
        x360.Extractor(0, 0, 64, 90, np.eye(3)[None])
    with pytest.raises(x360.X360Error):
        x360.Extractor(64, 128, 32, 200.0, np.eye(3)[None])


def test_no_cpu_fallback_without_device(x360):
    """On a box without CUDA every compute entry must fail loudly (never silently compute on the CPU)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    with pytest.raises(x360.X360Error, match="no CUDA device|CUDA"):
        x360.ImageUtils.calculate_blur_score(np.zeros((8, 8), np.uint8))
    with pytest.raises(x360.X360Error):
        x360.Extractor(64, 128, 32, 90, np.eye(3)[None])
    with pytest.raises(x360.X360Error):
        x360.remap(np.zeros((8, 8, 3), np.uint8), np.zeros((4, 4), np.float32), np.zeros((4, 4), np.float32))
    assert x360.ImageUtils.calculate_blur_score(None) == 0.0     # image_utils.py:18-19: never raises


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "360extractor_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"(from|import)\s+oracle|oracle[/.]|libx360_oracle", text), f"{f} reaches into oracle/"
                assert "import cv2" not in text, f"{f} imports cv2"


# ---- blur accept / reject machine (processor.py:341-376) --------------------------------------
def _reference_fsm(scores, enabled, smart, threshold):
    """Independent transliteration of the decision order for the test (NOT imported by the product)."""
    from collections import deque
    hist, skips, keep = deque(maxlen=10), 0, []
    for s in scores:
        if not enabled:
            keep.append(True)
            continue
        blurry = False
        if smart:
            if s < threshold:
                blurry = True
            elif len(hist) > 0 and s < (sum(hist) / len(hist)) * 0.6:
                blurry = True
            if blurry:
                skips += 1
                if skips > 5:
                    blurry, skips = False, 0
            if not blurry:
                skips = 0
                hist.append(s)
        elif s < threshold:
            blurry = True
        keep.append(not blurry)
    return keep


@pytest.mark.parametrize("smart", [False, True])
def test_blur_filter_replay(x360, smart):
    rng = np.random.default_rng(5)
    scores = np.concatenate([rng.uniform(50, 400, 40), rng.uniform(0, 90, 12), rng.uniform(300, 900, 20),
                             rng.uniform(100, 160, 30)])
    table = scores.reshape(-1, 6)
    f = x360.BlurFilter(True, smart, 100.0)
    got = f.replay(table)
    want = np.array(_reference_fsm(list(scores), True, smart, 100.0)).reshape(-1, 6)
    assert np.array_equal(np.array(got), want)
    assert f.skipped == int((~want).sum())
    if smart:
        assert f.forced >= 1                      # 12 sub-threshold scores in a row force an accept
    off = x360.BlurFilter(False, smart, 100.0)
    assert all(all(r) for r in off.replay(table))


def test_blur_filter_from_settings_defaults(x360):
    f = x360.BlurFilter.from_settings({})
    assert (f.enabled, f.smart, f.threshold) == (False, False, 100.0)   # processor.py:219-221
    f = x360.BlurFilter.from_settings({"blur_filter_enabled": True, "smart_blur_enabled": True, "blur_threshold": 42.5})
    assert (f.enabled, f.smart, f.threshold) == (True, True, 42.5)


# ---- plan preview: view units that share a map, staging capacity (host-only logic of the tiled path) ----
def test_plan_preview_view_units(x360):
    """Views that differ only by a yaw (ring: geometry.py:71-75; the 4 horizontal cube faces: :30-35)
    share one map evaluation; their map_x is shifted by yaw / 360 * W."""
    def preview(layout, n, pitch, H=2880, W=5760, d=1600):
        views = x360.GeometryProcessor.generate_views(n, pitch_offset=pitch, layout_mode=layout)
        return views, x360.plan_preview(H, W, d, 90, x360.rotation_stack(views)[1])

    views, pv = preview("ring", 8, 0)
    assert pv["n_units"] == 1 and pv["unit_of_view"] == [0] * 8
    np.testing.assert_allclose(pv["shift_px"], [v[1] / 360.0 * 5760 for v in views], rtol=0, atol=1e-9)
    views, pv = preview("ring", 8, -20)                                  # inclination does not break the sharing
    assert pv["n_units"] == 1
    np.testing.assert_allclose(pv["shift_px"], [v[1] / 360.0 * 5760 for v in views], rtol=0, atol=1e-9)
    views, pv = preview("cube", 6, 0)                                    # Front/Right/Back/Left share, Up and Down do not
    assert pv["n_units"] == 3 and pv["unit_of_view"] == [0, 0, 0, 0, 1, 2]
    np.testing.assert_allclose(pv["shift_px"][:4], [0, 1440, 2880, 4320], rtol=0, atol=1e-9)
    views, pv = preview("fibonacci", 7, -20)                             # every view has its own pitch
    assert pv["n_units"] == 7 and pv["shift_px"] == [0.0] * 7
    views, pv = preview("ring", 36, 0)                                   # at most 8 views per unit
    assert pv["n_units"] == 5 and pv["unit_of_view"].count(0) == 8 and pv["unit_of_view"].count(4) == 4
    assert all(0.0 <= s < 5760 for s in pv["shift_px"])


def test_plan_preview_pole_pixel_views_stay_alone(x360):
    # pitch 45 + fov 90 / 2 = 90 deg, even d: pixel (d / 2, 0) of every view looks straight at the pole -> no map sharing
    views = x360.GeometryProcessor.generate_views(8, 45, "ring")
    R = x360.rotation_stack(views)[1]
    assert x360.plan_preview(512, 1024, 228, 90, R)["n_units"] == 8
    assert x360.plan_preview(512, 1024, 229, 90, R)["n_units"] == 1           # odd d: the pole falls between pixels
    assert x360.plan_preview(512, 1024, 228, 80, R)["n_units"] == 1           # pole outside the view


def test_plan_preview_capacity(x360):
    views = x360.GeometryProcessor.generate_views(8, 0, "ring")
    pv = x360.plan_preview(2880, 5760, 1600, 90, x360.rotation_stack(views)[1])
    # 32 output rows cover at most 32 * (W / 2 pi) / f = 36.7 source rows (+ the lower tap, + slack)
    assert 38 <= pv["rows_max"] <= 48 and pv["rows_max"] % 8 == 0
    assert pv["bw_max"] in (96, 128, 144, 192, 240, 256) and pv["tile_h"] == 32      # 64-byte steps on the raw-gather plans
    # a strongly minifying geometry (11520-wide source, 2048^2 views: 1.8 source px per output px) works in
    # 16-row tiles so that the footprint buffers still leave several CTAs per SM
    views12 = x360.GeometryProcessor.generate_views(12, 0, "ring")
    pv12 = x360.plan_preview(5760, 11520, 2048, 90, x360.rotation_stack(views12, [0, 3, 6, 9])[1])
    assert pv12["with_scores"]["tile_h"] == 16                           # pair-record planes (fused score)
    # without the score: 16-row tiles on the planes, or — if the three raw buffers leave enough CTAs per SM — the raw gather
    assert pv12["tile_h"] == 16 or (pv12["tile_h"] == 32 and pv12["bw_max"] % 64 == 0)
    with pytest.raises(x360.X360Error):
        x360.plan_preview(2880, 5760, 1600, 190, x360.rotation_stack(views)[1])


def test_column_only_map_x(x360, oracle):
    """map_x is independent of the output row exactly for yaw-only rotations (R01 = R21 = 0): ring layouts at pitch 0
    and the horizontal cube faces; any pitch, the Up / Down faces or a roll bring yn back into v0 / v2."""
    def R_of(layout, n, pitch, active=None):
        views = x360.GeometryProcessor.generate_views(n, pitch_offset=pitch, layout_mode=layout)
        return x360.rotation_stack(views, active)[1]
    assert x360.column_only_map_x(R_of("ring", 8, 0))
    assert x360.column_only_map_x(R_of("ring", 12, 0))
    assert not x360.column_only_map_x(R_of("ring", 8, -20))
    assert not x360.column_only_map_x(R_of("cube", 6, 0))
    assert x360.column_only_map_x(R_of("cube", 6, 0, active=[0, 1, 2, 3]))
    assert not x360.column_only_map_x(R_of("fibonacci", 8, 0))
    assert not x360.column_only_map_x(x360.GeometryProcessor.get_rotation_matrix(30.0, 0.0, 5.0)[None])
    # numerically: for such R the reference map_x has identical rows
    R = x360.GeometryProcessor.get_rotation_matrix(135.0, 0.0, 0.0)
    mx, _ = oracle.make_map(180, 360, 48, 48, 90, R)
    assert x360.column_only_map_x(R[None]) and np.all(mx == mx[0:1])


# ---- device hand-off to the masker: host-side selection logic -------------------------------
def test_mask_face_selection_mirrors_reference(x360):
    cases = [None, [], "", "Front, back", ["Up", " down "], "  ,", ["FRONT"], "left"]
    want = [None, None, None, {"front", "back"}, {"up", "down"}, None, {"front"}, {"left"}]
    for c, w in zip(cases, want):
        assert x360.normalize_mask_faces(c) == w
    ref_src = "/root/reference/src"
    if os.path.exists(os.path.join(ref_src, "core", "settings_manager.py")):        # the reference itself, where it exists
        sys.path.insert(0, ref_src)
        try:
            from core.settings_manager import normalize_mask_faces as ref_norm
            for c in cases:
                assert x360.normalize_mask_faces(c) == ref_norm(c)
        finally:
            sys.path.remove(ref_src)
    names = ["Front", "Right", "Back", "Left", "Up", "Down"]
    assert x360.eligible_indices(names, None) == [0, 1, 2, 3, 4, 5]
    assert x360.eligible_indices(names, {"front", "down"}) == [0, 5]             # processor.py:405-408, case-insensitive
    assert x360.eligible_indices(names, {"nothing"}) == []
    # 2 frames x 6 views, only Up / Down masked, frame 1's Up view rejected by the blur filter
    keep = np.ones((2, 6), bool)
    keep[1, 4] = False
    assert x360.masker_slots(2, names, "up,down") == [4, 5, 10, 11]
    assert x360.masker_slots(2, names, "up,down", keep) == [4, 5, 11]
    assert x360.masker_slots(1, names) == [0, 1, 2, 3, 4, 5]
    res = x360.scatter_results([4, 5], [("imgU", "maskU"), ("imgD", "maskD")], [f"img{i}" for i in range(6)])
    assert res[0] == ("img0", None) and res[4] == ("imgU", "maskU") and res[5] == ("imgD", "maskD") and len(res) == 6


# ---- the caller contract pinned to the reference's own loop (tests/golden/make_worker_golden.py) ----------------------
def _worker_cases():
    import json
    with open(os.path.join(GOLDEN, "worker_cases.json")) as f:
        return json.load(f)


def worker_expectations(x360, case):
    """(names of the emitted views, reference scores [F][V] or None, kept (frame, camera) pairs) of one recorded scenario."""
    st = case["settings"]
    layout = "ring" if st["layout_mode"] == "adaptive" else st["layout_mode"]
    views = x360.GeometryProcessor.generate_views(st["camera_count"], pitch_offset=st["pitch_offset"], layout_mode=layout)
    names, _ = x360.rotation_stack(views, st.get("active_cameras"))
    scores = [float.fromhex(h) for h in case["scores_in_call_order"]]
    table = np.array(scores).reshape(-1, len(names)) if scores else None
    kept = []
    for fn in case["kept"]:                                    # clip_frame000003_View_2.png (processor.py:446)
        stem = fn[: -len(".png")]
        frame, cam = stem[len("clip_frame"):].split("_", 1)
        kept.append((int(frame), cam))
    order = {n: i for i, n in enumerate(names)}
    kept.sort(key=lambda fc: (fc[0], order[fc[1]]))
    return names, table, kept


@pytest.mark.parametrize("name", ["smart_ring4", "plain_threshold_cube", "smart_active_cameras", "filter_off_fibonacci",
                                  "smart_sharpened_lanczos", "legacy_adaptive_layout"])
def test_blur_machine_reproduces_the_reference_worker(x360, name):
    """BlurFilter.replay over the scores the REAL ProcessingWorker computed (recorded in call order) keeps exactly the files
    the worker wrote: threshold rejections, rolling-average rejections and forced accepts (processor.py:341-376)."""
    case = _worker_cases()["scenarios"][name]
    names, table, kept = worker_expectations(x360, case)
    filt = x360.BlurFilter.from_settings(case["settings"])
    n_frames = _worker_cases()["clip"]["frames"]
    if table is None:
        assert not filt.enabled
        want = [(f, n) for f in range(n_frames) for n in names]
    else:
        assert table.shape == (n_frames, len(names))            # one score per (frame, emitted view), in loop order
        keep = filt.replay(table)
        want = [(f, names[v]) for f in range(n_frames) for v in range(len(names)) if keep[f][v]]
    assert want == kept
    if "smart" in name or "legacy" in name:
        assert filt.forced >= 1 and filt.skipped >= 1


def test_jpeg_header_matches_the_oracle(x360):
    """Host-only piece of the device JPEG encoder: the SOI..SOS header (quantisation tables for the quality, SOF0 4:2:0,
    Annex-K Huffman tables) equals the oracle's, which is pinned to cv2's files."""
    from oracle import c_oracle
    for (h, w, q) in [(1600, 1600, 95), (48, 64, 75), (1, 1, 100), (2048, 2048, 1), (100, 333, 50)]:
        assert x360.jpeg_header(h, w, q) == c_oracle.jpeg_header(h, w, q), (h, w, q)
    assert len(x360.jpeg_header(16, 16)) == 623


def test_lanczos4_factors_rebuild_the_table(x360):
    """The tiled Lanczos4 kernel computes its 64 weights per pixel from the table's factors instead of loading a table row:
    float32 products of the 1-D coefficients, x 2^15, round half to even, saturate, plus the one-tap fix-up.  Rebuilt here
    with numpy float32 arithmetic, the factors give exactly the table (which the oracle pins to cv2)."""
    from oracle import c_oracle
    L = x360._capi.lib()
    k = np.zeros((32, 8), np.float32)
    fix = np.zeros(1024, np.uint32)
    x360._capi.check(L.x360_lanczos4_factors(k.ctypes.data, fix.ctypes.data))
    tab = np.zeros((32, 32, 8, 8), np.int16)
    x360._capi.check(L.x360_lanczos4_table(tab.ctypes.data))
    p = (k[:, None, :, None] * k[None, :, None, :]).astype(np.float32)              # [fy][fx][r][c], float32 products
    magic = np.float32(12582912.0)
    bits = (p * np.float32(32768.0) + magic).astype(np.float32).view(np.uint32)      # the kernel's FFMA(p, 2^15, 1.5 * 2^23): p * 2^15 is exact
    v = (bits.astype(np.int64) - 0x4B400000)
    v = np.minimum(v, 32767)                                                         # only 1.0 * 1.0 saturates
    pos = (fix & 3).reshape(32, 32)
    delta = (fix.view(np.int32) >> 16).reshape(32, 32)
    for q in range(4):
        v[:, :, 4 + (q >> 1), 4 + (q & 1)] += np.where(pos == q, delta, 0)
    assert np.array_equal(v.astype(np.int16), tab)
    assert np.array_equal(tab.reshape(1024, 8, 8), c_oracle.lanczos4_table().reshape(1024, 8, 8))
    assert (v.reshape(1024, 64).sum(axis=1) == 32768).all()                          # what the fix-up is for
    assert np.count_nonzero(delta) == 856                                            # 856 of 1024 entries are fixed up
    # the kernel fuses the product and the rounding into ONE FFMA(k[fy][r], k[fx][c] * 2^15, 1.5 * 2^23): a single rounding of
    # the exact product instead of OpenCV's two (to float32, then to the integer) — the same integer for every entry of the table
    exact = k[:, None, :, None].astype(np.float64) * k[None, :, None, :].astype(np.float64) * 32768.0     # exact in fp64 (24 + 24 bits)
    assert np.array_equal(np.rint(exact), np.rint(p.astype(np.float64) * 32768.0))


def test_c_blur_machine_equals_the_python_replay(x360):
    """x360_blur_accept (the machine inside x360_extract_host_filtered) makes the same decisions as BlurFilter — which
    test_blur_machine_reproduces_the_reference_worker pins to the reference's loop — on the recorded scores and on random ones."""
    cap = x360._capi
    L = cap.lib()

    def run_c(scores, enabled, smart, th):
        st = cap.BlurState()
        L.x360_blur_state_init(st, enabled, smart, th)
        return [bool(L.x360_blur_accept(st, float(s))) for s in scores], int(st.skipped), int(st.forced)

    for name, case in _worker_cases()["scenarios"].items():
        scores = [float.fromhex(h) for h in case["scores_in_call_order"]]
        f = x360.BlurFilter.from_settings(case["settings"])
        want = [f.accept(s) for s in scores]
        assert run_c(scores, int(f.enabled), int(f.smart), float(f.threshold)) == (want, f.skipped, f.forced), name
    rng = np.random.default_rng(0)
    for _ in range(100):
        scores = np.concatenate([rng.uniform(0, 500, rng.integers(1, 60)), rng.uniform(0, 30, rng.integers(0, 20)), rng.uniform(100, 900, rng.integers(0, 40))])
        rng.shuffle(scores)
        for smart in (0, 1):
            th = float(rng.uniform(10, 300))
            f = x360.BlurFilter(True, bool(smart), th)
            want = [f.accept(float(s)) for s in scores]
            assert run_c(scores, 1, smart, th) == (want, f.skipped, f.forced)
            # the table form the sharded bench uses (x360_blur_replay): the same flags, and a state that carries on
            table = scores[: len(scores) // 3 * 3].reshape(-1, 3)
            f2 = x360.BlurFilter(True, bool(smart), th)
            assert np.array_equal(x360.blur_replay(table, smart=bool(smart), threshold=th), np.array(f2.replay(table), dtype=bool).reshape(table.shape))
    st = cap.BlurState()
    L.x360_blur_state_init(st, 1, 1, 100.0)
    f3 = x360.BlurFilter(True, True, 100.0)
    for part in np.array_split(rng.uniform(0, 400, 90), 4):
        assert np.array_equal(x360.blur_replay(part, state=st), np.array([f3.accept(float(s)) for s in part]))
    assert (int(st.skipped), int(st.forced)) == (f3.skipped, f3.forced)
