"""Randomised parity run on a B200: seeded random geometries (sizes, fov, layout, pitch, frames per call, scores on / off,
both interpolations, plan flags) through the C-ABI against the C oracle, bit for bit.

    python scripts/fuzz_parity.py report.json [n_cases] [seed]

`run()` is also what tests/test_gpu_parity.py::test_fuzz_seed_against_oracle calls with a small case count."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def run(n_cases=200, seed=2024, verbose=True):
    import x360
    from oracle import c_oracle
    rng = np.random.default_rng(seed)
    report = {"lib": x360.version(), "seed": seed, "cases": 0, "failures": [], "pixels": 0, "by_mode": {}, "by_flags": {}}
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
        # plan variants: the plan's own choice, or an explicit override (x360_params.flags)
        flags = int(rng.choice([0, 0, 0, x360.FLAG_NO_WIDE, x360.FLAG_PLANES, x360.FLAG_FUSED_SCORE, x360.FLAG_SPLIT_SCORE,
                                x360.FLAG_TILE_H16, x360.FLAG_TILE_H32 | x360.FLAG_FUSED_SCORE]))
        if rng.random() < 0.25:                            # a wide source under a small view: seam boxes, wide boxes, the second launch
            H = int(rng.integers(200, 420)) * 2
            W = 2 * H
            d = int(rng.integers(64, 200))
            F = min(F, 3)
        if rng.random() < 0.3:                             # whole 32 x 32 blocks: general views may take transposed tiles
            d = max(32, (d + 31) // 32 * 32)
        frames = rng.integers(0, 256, (F, H, W, 3), dtype=np.uint8)
        views = x360.GeometryProcessor.generate_views(n, pitch_offset=pitch, layout_mode=layout)
        names, R = x360.rotation_stack(views)
        if len(names) > 8:
            R = R[:8]
        want, ws, ws2 = c_oracle.extract(frames, R, d, fov, interp, want_scores=scores)
        desc = dict(case=c, H=H, W=W, d=d, fov=fov, layout=layout, n=n, pitch=pitch, F=F, interp=interp, scores=scores, flags=flags)
        try:
            with x360.Extractor(H, W, d, fov, R, "lanczos" if interp else "linear", max_batch=int(rng.integers(1, F + 1)), flags=flags) as ex:
                if scores:
                    got, s, s2 = ex.extract_sums(frames)
                    ok = np.array_equal(got, want) and np.array_equal(s, ws) and np.array_equal(s2, ws2)
                else:
                    got, _ = ex.extract(frames, want_scores=False)
                    ok = np.array_equal(got, want)
                modes = ex.tile_modes()
                modes["transposed_blocks"] = ex.transposed_blocks()
        except Exception as e:                            # noqa: BLE001 - the report carries the message
            ok, modes = False, {"error": str(e)}
        report["cases"] += 1
        report["pixels"] += int(want.size // 3)
        report["by_flags"][str(flags)] = report["by_flags"].get(str(flags), 0) + 1
        for k, v in modes.items():
            if isinstance(v, int):
                report["by_mode"][k] = report["by_mode"].get(k, 0) + v
        if not ok:
            desc["modes"] = modes
            if "error" not in modes:
                desc["differing_bytes"] = int(np.count_nonzero(got != want))
            report["failures"].append(desc)
            if verbose:
                print("FAIL", desc, flush=True)
    report["seconds"] = round(time.time() - t0, 1)
    report["all_exact"] = not report["failures"]
    return report


if __name__ == "__main__":
    rep = run(int(sys.argv[2]) if len(sys.argv) > 2 else 200, int(sys.argv[3]) if len(sys.argv) > 3 else 2024)
    json.dump(rep, open(sys.argv[1], "w"), indent=1)
    print({k: rep[k] for k in ("cases", "pixels", "by_mode", "by_flags", "seconds", "all_exact")})
