#!/usr/bin/env python
"""Generate the committed golden fixtures from the REAL reference.

Runs only in the build container: imports ``core.geometry`` and
``utils.image_utils`` straight from ``/root/reference/src`` (unmodified) and
calls ``cv2.remap`` exactly as ``src/core/processor.py:334`` does.  The outputs
are what the oracle (``oracle/x360_oracle.c``) and the CUDA path are pinned to.

    python tests/golden/make_golden.py            # rewrites tests/golden/*.json, *.npz

Nothing under tests/ reads /root/reference at test time; only this script does.
"""
import hashlib
import json
import os
import sys

import numpy as np

REF = os.environ.get("X360_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(REF, "src"))

import cv2  # noqa: E402
from core.geometry import GeometryProcessor  # noqa: E402  (the reference's)
from utils.image_utils import ImageUtils  # noqa: E402  (the reference's)

HERE = os.path.dirname(os.path.abspath(__file__))


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def ref_remap(frame, mx, my, mode):
    flag = cv2.INTER_LANCZOS4 if mode == "lanczos" else cv2.INTER_LINEAR  # processor.py:199-200
    return cv2.remap(frame, mx, my, flag, borderMode=cv2.BORDER_WRAP)     # processor.py:334


def geometry_fixture():
    views = {}
    for layout, n, pitch in [("ring", 6, 0), ("ring", 8, -20), ("ring", 12, 0), ("ring", 7, 45),
                             ("ring", 2, 0), ("ring", 36, -45),
                             ("cube", 6, 0), ("cube", 12, -20), ("cube", 6, 45),
                             ("fibonacci", 4, 0), ("fibonacci", 8, 0), ("fibonacci", 16, 20),
                             ("fibonacci", 20, -20), ("fibonacci", 36, -90),
                             ("adaptive", 5, 0), ("nonsense", 3, -20)]:
        vs = GeometryProcessor.generate_views(n, pitch_offset=pitch, layout_mode=layout)
        views[f"{layout},{n},{pitch}"] = [[nm, float(y).hex(), float(p).hex(), int(r)] for nm, y, p, r in vs]
    rots = {}
    for ypr in [(0, 0, 0), (180, 0, 0), (0, 90, 0), (0, -90, 0), (90, -20, 0), (270, 45, 0),
                (312.49223594996215, 38.21166938294838, 0), (45, -20, 30), (123.4, -63.2, -17.5),
                (360, 0, 0), (60, 110, 0)]:
        R = GeometryProcessor.get_rotation_matrix(*ypr)
        rots[",".join(float(a).hex() for a in ypr)] = [float(v).hex() for v in R.reshape(-1)]
    return {"views": views, "rotations": rots}


def kat_fixture():
    """512x1024 noise, d=256, three layouts."""
    src = np.random.default_rng(12345).integers(0, 256, (512, 1024, 3), dtype=np.uint8)
    out = {"src_sha": sha(src), "d": 256, "cases": {}}
    for layout, n, pitch, fov in [("cube", 6, 0, 90), ("ring", 8, -20, 90), ("fibonacci", 20, -20, 110)]:
        views = GeometryProcessor.generate_views(n, pitch_offset=pitch, layout_mode=layout)
        maps, lin, lan, s_lin, s_lan, rs = [], [], [], [], [], []
        for name, y, p, r in views:
            mx, my = GeometryProcessor.create_rectilinear_map(512, 1024, 256, 256, fov, y, p, r)
            maps += [mx, my]
            a = ref_remap(src, mx, my, "linear")
            b = ref_remap(src, mx, my, "lanczos")
            lin.append(a)
            lan.append(b)
            s_lin.append(float(ImageUtils.calculate_blur_score(a)))
            s_lan.append(float(ImageUtils.calculate_blur_score(b)))
            rs.append([float(v).hex() for v in GeometryProcessor.get_rotation_matrix(y, p, r).reshape(-1)])
        out["cases"][f"{layout},{n},{pitch},{fov}"] = {
            "R": rs,
            "maps_sha": sha(*maps), "linear_sha": sha(*lin), "lanczos_sha": sha(*lan),
            "linear_view_sha": [sha(a) for a in lin], "lanczos_view_sha": [sha(a) for a in lan],
            "linear_scores": [s.hex() for s in s_lin], "lanczos_scores": [s.hex() for s in s_lan],
        }
    return out


def small_fixture():
    """Raw bytes for small cases (poles, seam, roll, odd sizes, non-square src ratio)."""
    rng = np.random.default_rng(7)
    src = rng.integers(0, 256, (96, 192, 3), dtype=np.uint8)
    arrays = {"src": src}
    meta = []
    cases = [  # (d, fov, yaw, pitch, roll)
        (48, 90, 0, 0, 0), (48, 90, 180, 0, 0), (48, 90, 0, 90, 0), (48, 90, 0, -90, 0),
        (48, 120, 90, -20, 0), (48, 60, 270, 45, 0), (47, 110, 312.49223594996215, 38.21166938294838, 0),
        (33, 100, 45, -20, 30), (48, 90, 123.4, -63.2, -17.5), (16, 90, 60, 110, 0),
    ]
    for k, (d, fov, y, p, r) in enumerate(cases):
        mx, my = GeometryProcessor.create_rectilinear_map(96, 192, d, d, fov, y, p, r)
        a = ref_remap(src, mx, my, "linear")
        b = ref_remap(src, mx, my, "lanczos")
        arrays[f"mx{k}"], arrays[f"my{k}"] = mx, my
        arrays[f"lin{k}"], arrays[f"lan{k}"] = a, b
        meta.append({"d": d, "fov": fov, "yaw": float(y).hex(), "pitch": float(p).hex(), "roll": float(r).hex(),
                     "R": [float(v).hex() for v in GeometryProcessor.get_rotation_matrix(y, p, r).reshape(-1)],
                     "lin_score": float(ImageUtils.calculate_blur_score(a)).hex(),
                     "lan_score": float(ImageUtils.calculate_blur_score(b)).hex()})
    return arrays, meta


def blur_fixture():
    out = {}
    edge = np.zeros((100, 100), np.uint8)
    edge[:, 50:] = 255                                       # tests/test_core.py:87-88
    out["edge_100"] = float(ImageUtils.calculate_blur_score(edge)).hex()
    out["uniform_100"] = float(ImageUtils.calculate_blur_score(np.ones((100, 100), np.uint8) * 128)).hex()
    out["none"] = float(ImageUtils.calculate_blur_score(None)).hex()
    rng = np.random.default_rng(99)
    for name, shape in [("color_100", (100, 100, 3)), ("color_37x53", (37, 53, 3)), ("gray_64x17", (64, 17)),
                        ("color_2x2", (2, 2, 3)), ("color_1x5", (1, 5, 3)), ("color_5x1", (5, 1, 3))]:
        img = rng.integers(0, 256, shape, dtype=np.uint8)
        out[name] = {"sha": sha(img), "shape": list(shape), "seed_order": name,
                     "score": float(ImageUtils.calculate_blur_score(img)).hex()}
    return out


UNSHARP_CASES = [  # (name, shape, strength, kind)
    ("noise_64", (64, 64, 3), 0.5, "noise"), ("noise_37x53", (37, 53, 3), 0.3, "noise"),
    ("noise_5x5", (5, 5, 3), 1.7, "noise"), ("noise_1x20", (1, 20, 3), 0.5, "noise"),
    ("noise_20x1", (20, 1, 3), 0.9, "noise"), ("smooth_100x96", (100, 96, 3), 0.5, "smooth"),
    ("smooth_48x64", (48, 64, 3), 2.0, "smooth"), ("smooth_96x160", (96, 160, 3), 0.1, "smooth"),
    ("noise_70x80", (70, 80, 3), 1.3, "noise"), ("smooth_33x47", (33, 47, 3), 0.05, "smooth"),
]


def unsharp_image(k, shape, kind):
    """Seeded input of case k: uint8 noise, or the same generator at 1/8 size upsampled with INTER_CUBIC."""
    rng = np.random.default_rng(4000 + k)
    if kind == "noise":
        return rng.integers(0, 256, shape, dtype=np.uint8)
    small = rng.integers(0, 256, (max(2, shape[0] // 8), max(2, shape[1] // 8), shape[2]), dtype=np.uint8)
    return cv2.resize(small, (shape[1], shape[0]), interpolation=cv2.INTER_CUBIC)


def ref_unsharp(img, s):
    gaussian = cv2.GaussianBlur(img, (0, 0), 2.0)                      # processor.py:380
    return cv2.addWeighted(img, 1.0 + s, gaussian, -s, 0)              # processor.py:381


def unsharp_fixture():
    """processor.py:379-381 on seeded images; inputs and outputs are committed (small), plus the blur alone
    and the full byte-pair table of addWeighted for ten strengths (hashes)."""
    arrays, meta = {}, []
    for k, (name, shape, s, kind) in enumerate(UNSHARP_CASES):
        img = unsharp_image(k, shape, kind)
        arrays[f"in{k}"] = img
        arrays[f"blur{k}"] = cv2.GaussianBlur(img, (0, 0), 2.0)
        arrays[f"out{k}"] = ref_unsharp(img, s)
        meta.append({"name": name, "shape": list(shape), "strength": float(s).hex(), "kind": kind})
    a, g = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), indexing="ij")
    a, g = np.ascontiguousarray(a), np.ascontiguousarray(g)
    table = {}
    for s in (0.5, 0.1, 0.3, 1.0, 2.0, 0.7, 1.3, 0.05, 1.9, 0.25):
        table[float(s).hex()] = sha(cv2.addWeighted(a, 1.0 + s, g, -s, 0))
    impulse = np.zeros((25, 25), np.uint8)
    impulse[12, :] = 255                                               # a row of 255s: the blur's column = round(255 k_j)
    col = cv2.GaussianBlur(impulse, (0, 0), 2.0)[6:19, 12]
    return arrays, {"cases": meta, "add_weighted_table_sha": table, "line_response": [int(v) for v in col]}


def jpeg_image(kind, h, w, seed):
    rng = np.random.default_rng(seed)
    if kind == "noise":
        return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    if kind == "smooth":
        lo = rng.integers(0, 256, (max(1, h // 8) + 1, max(1, w // 8) + 1, 3), dtype=np.uint8)
        return np.ascontiguousarray(cv2.resize(lo, (w, h), interpolation=cv2.INTER_CUBIC))
    if kind == "flat":
        return np.full((h, w, 3), int(rng.integers(0, 256)), np.uint8)
    return (rng.integers(0, 2, (h, w, 3)) * 255).astype(np.uint8)          # saturated


JPEG_CASES = [(kind, h, w, q) for q in (95, 75, 100, 30) for kind in ("noise", "smooth", "flat", "sat")
              for (h, w) in ((16, 16), (48, 64), (17, 33), (100, 100), (8, 8), (233, 77), (24, 40), (1, 1), (256, 256))]


def jpeg_fixture():
    """What the reference's save step writes: cv2.imencode('.jpg', img, [IMWRITE_JPEG_QUALITY, q]) == the bytes cv2.imwrite
    puts in the file (src/utils/file_manager.py:21-32, src/core/processor.py:121-123)."""
    out = []
    for i, (kind, h, w, q) in enumerate(JPEG_CASES):
        img = jpeg_image(kind, h, w, 1000 + i)
        ok, buf = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, q])
        assert ok
        out.append({"kind": kind, "h": h, "w": w, "quality": q, "seed": 1000 + i, "image_sha": sha(img), "bytes": int(buf.size),
                    "jpeg_sha": hashlib.sha256(buf.tobytes()).hexdigest()})
    return {"cases": out, "cv2": cv2.__version__, "libjpeg": "libjpeg-turbo (bundled with the opencv-python wheel)"}


def main():
    if "--only-jpeg" in sys.argv:
        with open(os.path.join(HERE, "jpeg_cases.json"), "w") as f:
            json.dump(jpeg_fixture(), f, indent=1)
        print("jpeg fixtures written")
        return
    if "--only-unsharp" in sys.argv:
        arrays, meta = unsharp_fixture()
        np.savez_compressed(os.path.join(HERE, "unsharp_cases.npz"), **arrays)
        with open(os.path.join(HERE, "unsharp.json"), "w") as f:
            json.dump(meta, f, indent=1)
        print("unsharp fixtures written")
        return
    with open(os.path.join(HERE, "geometry.json"), "w") as f:
        json.dump(geometry_fixture(), f, indent=1)
    with open(os.path.join(HERE, "kat_512x1024.json"), "w") as f:
        json.dump(kat_fixture(), f, indent=1)
    arrays, meta = small_fixture()
    np.savez_compressed(os.path.join(HERE, "small_cases.npz"), **arrays)
    with open(os.path.join(HERE, "small_cases.json"), "w") as f:
        json.dump(meta, f, indent=1)
    with open(os.path.join(HERE, "blur.json"), "w") as f:
        json.dump(blur_fixture(), f, indent=1)
    arrays, meta = unsharp_fixture()
    np.savez_compressed(os.path.join(HERE, "unsharp_cases.npz"), **arrays)
    with open(os.path.join(HERE, "unsharp.json"), "w") as f:
        json.dump(meta, f, indent=1)
    with open(os.path.join(HERE, "jpeg_cases.json"), "w") as f:
        json.dump(jpeg_fixture(), f, indent=1)
    info = {"cv2": cv2.__version__, "numpy": np.__version__, "reference": REF,
            "threads": cv2.getNumThreads()}
    with open(os.path.join(HERE, "PROVENANCE.json"), "w") as f:
        json.dump(info, f, indent=1)
    print("golden fixtures written:", info)


if __name__ == "__main__":
    main()
