"""Randomised parity run on a B200: seeded random geometries (sizes, fov, layout, pitch, frames per call, scores on / off,
both interpolations) through the C-ABI against the C oracle, bit for bit.  Writes a JSON report to argv[1];
argv[2] = number of cases (default 200), argv[3] = seed."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import x360
from oracle import c_oracle

n_cases = int(sys.argv[2]) if len(sys.argv) > 2 else 200
rng = np.random.default_rng(int(sys.argv[3]) if len(sys.argv) > 3 else 2024)
report = {"lib": x360.version(), "cases": 0, "failures": [], "pixels": 0, "by_mode": {}}
t0 = time.time()
for c in range(n_cases):
    H = int(rng.integers(48, 520))
    W = int(rng.choice([2 * H, 2 * H + 2, int(rng.integers(96, 1100))]))
    if rng.random() < 0.6:
        W = (W + 15) // 16 * 16                       # TMA-eligible pitch
    d = int(rng.integers(17, 330))
    fov = float(rng.choice([60, 75, 90, 100, 110, 120, int(rng.integers(40, 140))]))
    layout = str(rng.choice(["ring", "ring", "cube", "fibonacci"]))
    n = int(rng.integers(1, 13))
    pitch = int(rng.choice([0, 0, 0, -20, 20, 45, -45, -90, int(rng.integers(-90, 91))]))
    if rng.random() < 0.15:                            # a view edge or centre exactly on a pole / the horizon
        pitch = int(rng.choice([90 - fov / 2, fov / 2 - 90, 90, -90, fov / 2, -fov / 2]))
        d += d & 1                                     # even size: pixel (d / 2, 0) lies on the optical axis' column
    F = int(rng.integers(1, 8))
    interp = 1 if rng.random() < 0.2 else 0
    scores = bool(rng.random() < 0.4)
    frames = rng.integers(0, 256, (F, H, W, 3), dtype=np.uint8)
    views = x360.GeometryProcessor.generate_views(n, pitch_offset=pitch, layout_mode=layout)
    names, R = x360.rotation_stack(views)
    if len(names) > 8:
        R = R[:8]
    want, ws, ws2 = c_oracle.extract(frames, R, d, fov, interp, want_scores=scores)
    desc = dict(case=c, H=H, W=W, d=d, fov=fov, layout=layout, n=n, pitch=pitch, F=F, interp=interp, scores=scores)
    try:
        with x360.Extractor(H, W, d, fov, R, "lanczos" if interp else "linear", max_batch=int(rng.integers(1, F + 1))) as ex:
            if scores:
                got, s, s2 = ex.extract_sums(frames)
                ok = np.array_equal(got, want) and np.array_equal(s, ws) and np.array_equal(s2, ws2)
            else:
                got, _ = ex.extract(frames, want_scores=False)
                ok = np.array_equal(got, want)
            modes = ex.tile_modes()
    except Exception as e:                            # noqa: BLE001 - the report carries the message
        ok, modes = False, {"error": str(e)}
    report["cases"] += 1
    report["pixels"] += int(want.size // 3)
    for k, v in modes.items():
        if isinstance(v, int):
            report["by_mode"][k] = report["by_mode"].get(k, 0) + v
    if not ok:
        desc["modes"] = modes
        if "error" not in modes:
            desc["differing_bytes"] = int(np.count_nonzero(got != want))
        report["failures"].append(desc)
        print("FAIL", desc, flush=True)
report["seconds"] = round(time.time() - t0, 1)
report["all_exact"] = not report["failures"]
json.dump(report, open(sys.argv[1], "w"), indent=1)
print({k: report[k] for k in ("cases", "pixels", "by_mode", "seconds", "all_exact")})
