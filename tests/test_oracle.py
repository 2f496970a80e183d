"""The oracle against the committed golden fixtures (generated from the REAL reference by
tests/golden/make_golden.py).  CPU only."""
import hashlib
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_json, unhex


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def test_known_answer_hashes_match_survey_appendix_b():
    """SURVEY.md Appendix B pins the first 16 hex digits of these hashes independently."""
    k = load_json("kat_512x1024.json")
    assert k["src_sha"].startswith("67fdcbc84f671703")
    c = k["cases"]
    assert c["cube,6,0,90"]["linear_sha"].startswith("68e4c77ebc05934e")
    assert c["cube,6,0,90"]["lanczos_sha"].startswith("313d38495c6c4d8f")
    assert c["ring,8,-20,90"]["maps_sha"].startswith("fc0c7e4b849f0e35")
    assert c["fibonacci,20,-20,110"]["lanczos_sha"].startswith("b38406a628fb73f6")
    assert unhex(c["cube,6,0,90"]["linear_scores"][0]) == 11793.479548677802


@pytest.mark.parametrize("case", ["cube,6,0,90", "ring,8,-20,90", "fibonacci,20,-20,110"])
def test_oracle_bit_exact_on_kat(oracle, case):
    k = load_json("kat_512x1024.json")
    src = np.random.default_rng(12345).integers(0, 256, (512, 1024, 3), dtype=np.uint8)
    assert sha(src) == k["src_sha"]
    g = k["cases"][case]
    fov = int(case.split(",")[3])
    maps, lin, lan = [], [], []
    for i, Rhex in enumerate(g["R"]):
        R = np.array([unhex(v) for v in Rhex]).reshape(3, 3)
        mx, my = oracle.make_map(512, 1024, 256, 256, fov, R)
        maps += [mx, my]
        a = oracle.remap(src, mx, my, oracle.INTERP_LINEAR)
        b = oracle.remap(src, mx, my, oracle.INTERP_LANCZOS4)
        assert sha(a) == g["linear_view_sha"][i]
        assert sha(b) == g["lanczos_view_sha"][i]
        # scores: exact integer sums -> variance; reference is cv2.Laplacian(...).var() in fp64
        assert oracle.blur_score(a) == pytest.approx(unhex(g["linear_scores"][i]), rel=1e-12)
        assert oracle.blur_score(b) == pytest.approx(unhex(g["lanczos_scores"][i]), rel=1e-12)
        lin.append(a)
        lan.append(b)
    assert sha(*maps) == g["maps_sha"]
    assert sha(*lin) == g["linear_sha"]
    assert sha(*lan) == g["lanczos_sha"]


def test_oracle_bit_exact_on_small_cases(oracle):
    """Raw-byte fixtures: poles, the +-pi seam, roll, odd output sizes, pitch past the pole."""
    z = np.load(GOLDEN + "/small_cases.npz")
    meta = load_json("small_cases.json")
    src = z["src"]
    for k, m in enumerate(meta):
        R = np.array([unhex(v) for v in m["R"]]).reshape(3, 3)
        mx, my = oracle.make_map(src.shape[0], src.shape[1], m["d"], m["d"], m["fov"], R)
        assert np.array_equal(mx, z[f"mx{k}"]) and np.array_equal(my, z[f"my{k}"]), f"maps of case {k}"
        assert np.array_equal(oracle.remap(src, mx, my, 0), z[f"lin{k}"]), f"linear case {k}"
        assert np.array_equal(oracle.remap(src, mx, my, 1), z[f"lan{k}"]), f"lanczos case {k}"
        assert oracle.blur_score(z[f"lin{k}"]) == pytest.approx(unhex(m["lin_score"]), rel=1e-12)
        assert oracle.blur_score(z[f"lan{k}"]) == pytest.approx(unhex(m["lan_score"]), rel=1e-12)


def test_oracle_extract_composition(oracle):
    """orc_extract (the batched entry the CUDA library replaces) == per-view calls."""
    z = np.load(GOLDEN + "/small_cases.npz")
    meta = load_json("small_cases.json")
    src = z["src"]
    sel = [k for k, m in enumerate(meta) if m["d"] == 48 and m["fov"] == 90]
    R = np.array([[unhex(v) for v in meta[k]["R"]] for k in sel])
    frames = np.stack([src, np.roll(src, 37, axis=1)])
    for interp, key in ((0, "lin"), (1, "lan")):
        out, s, s2 = oracle.extract(frames, R, 48, 90, interp)
        for j, k in enumerate(sel):
            assert np.array_equal(out[0, j], z[f"{key}{k}"])
            assert oracle.score_from_sums(s[0, j], s2[0, j], 48 * 48) == pytest.approx(unhex(meta[k][f"{key}_score"]), rel=1e-12)
        assert not np.array_equal(out[0], out[1])


def test_oracle_blur_known_answers(oracle):
    b = load_json("blur.json")
    edge = np.zeros((100, 100), np.uint8)
    edge[:, 50:] = 255
    assert oracle.blur_score(edge) == unhex(b["edge_100"]) == 1300.5
    assert oracle.blur_score(np.ones((100, 100), np.uint8) * 128) == unhex(b["uniform_100"]) == 0.0
    assert oracle.blur_score(None) == 0.0
    rng = np.random.default_rng(99)
    for name in ("color_100", "color_37x53", "gray_64x17", "color_2x2", "color_1x5", "color_5x1"):
        img = rng.integers(0, 256, b[name]["shape"], dtype=np.uint8)
        assert sha(img) == b[name]["sha"]
        assert oracle.blur_score(img) == pytest.approx(unhex(b[name]["score"]), rel=1e-12, abs=1e-9)


def test_cv2_path_matches_c_oracle(oracle):
    """The cv2 call-sequence restatement (CPU baseline) and the plain-C restatement agree bit for bit."""
    cv2_path = pytest.importorskip("oracle.cv2_path")
    if not cv2_path.have_cv2():
        pytest.skip("cv2 not installed")
    src = np.random.default_rng(3).integers(0, 256, (160, 320, 3), dtype=np.uint8)
    views = cv2_path.layout_views(5, -20, "fibonacci")
    maps = cv2_path.build_maps(160, 320, 72, 100, views)
    for mode, interp in (("linear", 0), ("lanczos", 1)):
        outs, scores = cv2_path.extract([src], views, maps, mode, want_scores=True)
        for j, (name, y, p, r) in enumerate(views):
            mx, my = oracle.make_map(160, 320, 72, 72, 100, cv2_path.rotation(y, p, r))
            assert np.array_equal(mx, maps[name][0]) and np.array_equal(my, maps[name][1])
            got = oracle.remap(src, mx, my, interp)
            assert np.array_equal(got, outs[0][j])
            assert oracle.blur_score(got) == pytest.approx(scores[0, j], rel=1e-12)


# ---- unsharp mask (processor.py:379-381): GaussianBlur(sigma 2) + addWeighted ------------------------------

def test_oracle_gauss_taps_and_add_weighted_table(oracle):
    u = load_json("unsharp.json")
    taps = oracle.gauss13_taps(2.0)
    assert taps == [1, 2, 7, 16, 31, 45, 52, 45, 31, 16, 7, 2, 1] and sum(taps) == 256
    assert taps == u["line_response"]          # cv2's response to a line of 255s is the taps themselves
    a, g = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), indexing="ij")
    for s_hex, want in u["add_weighted_table_sha"].items():
        assert sha(oracle.add_weighted(a, g, unhex(s_hex))) == want, f"strength {unhex(s_hex)}"


def test_oracle_unsharp_bit_exact_on_golden(oracle):
    u = load_json("unsharp.json")
    z = np.load(os.path.join(GOLDEN, "unsharp_cases.npz"))
    for k, m in enumerate(u["cases"]):
        img = z[f"in{k}"]
        assert list(img.shape) == m["shape"]
        s = unhex(m["strength"])
        assert np.array_equal(oracle.unsharp(img, 0.0), img), "strength 0 is the identity"
        # blur alone: addWeighted(img, 0, blur, 1) == blur  <=>  strength -1
        assert np.array_equal(oracle.unsharp(img, -1.0), z[f"blur{k}"]), m["name"]
        assert np.array_equal(oracle.unsharp(img, s), z[f"out{k}"]), m["name"]


# ---- baseline JPEG: the reference's save step (cv2.imwrite with IMWRITE_JPEG_QUALITY) ---------------------------------
def _jpeg_image(kind, h, w, seed):
    """Same generator as tests/golden/make_golden.py::jpeg_image (the fixture carries the image hash)."""
    import cv2
    rng = np.random.default_rng(seed)
    if kind == "noise":
        return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    if kind == "smooth":
        lo = rng.integers(0, 256, (max(1, h // 8) + 1, max(1, w // 8) + 1, 3), dtype=np.uint8)
        return np.ascontiguousarray(cv2.resize(lo, (w, h), interpolation=cv2.INTER_CUBIC))
    if kind == "flat":
        return np.full((h, w, 3), int(rng.integers(0, 256)), np.uint8)
    return (rng.integers(0, 2, (h, w, 3)) * 255).astype(np.uint8)


def test_jpeg_oracle_matches_cv2_fixture(oracle):
    """orc_jpeg_encode == the bytes cv2 wrote when the fixture was made (sha256 per file), for every committed case:
    all qualities, sizes that are not multiples of 16 (dummy blocks, edge expansion), flat and saturated images."""
    import hashlib
    fx = load_json("jpeg_cases.json")
    assert len(fx["cases"]) >= 100
    for c in fx["cases"]:
        img = _jpeg_image(c["kind"], c["h"], c["w"], c["seed"])
        h = hashlib.sha256()
        h.update(np.ascontiguousarray(img).tobytes())
        if h.hexdigest() != c["image_sha"]:
            pytest.skip("this cv2 build resizes differently from the one that made the fixture: image hashes differ")
        got = oracle.jpeg_encode(img, c["quality"])
        assert len(got) == c["bytes"] and hashlib.sha256(got).hexdigest() == c["jpeg_sha"], c


def test_jpeg_oracle_matches_live_cv2(oracle):
    """The same against the cv2 of the box the test runs on (the wheel is part of the image)."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(99)
    for (h, w, q) in [(32, 32, 95), (50, 70, 95), (64, 48, 60), (31, 17, 90), (128, 256, 95), (40, 24, 100), (400, 400, 95)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        img[: h // 2] = (img[: h // 2].astype(np.uint16) // 4 + 100).astype(np.uint8)            # a calmer half
        ok, buf = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, q])
        assert ok and oracle.jpeg_encode(img, q) == buf.tobytes(), (h, w, q)
        # and it decodes to the same pixels as cv2's own file
        assert np.array_equal(cv2.imdecode(np.frombuffer(oracle.jpeg_encode(img, q), np.uint8), cv2.IMREAD_COLOR), cv2.imdecode(buf, cv2.IMREAD_COLOR))
