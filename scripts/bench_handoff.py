"""Kernel-only throughput of the masker hand-off (x360_views_to_nchw) on device-resident views.
Prints one JSON line: Mpix/s, ms per batch, algorithmic GB/s (3 B read + 12 B written per pixel) vs the measured HBM peak."""
import argparse, importlib, json, os, sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
x360 = importlib.import_module("360extractor_b200")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--views", type=int, default=8)
    ap.add_argument("--d", type=int, default=1600)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    g = torch.Generator(device="cuda").manual_seed(0)
    views = torch.randint(0, 256, (a.frames, a.views, a.d, a.d, 3), dtype=torch.uint8, device="cuda", generator=g)
    names = [f"View_{i}" for i in range(a.views)]
    out = torch.empty((a.frames * a.views, 3, a.d, a.d), dtype=torch.float32, device="cuda")
    for _ in range(a.warmup):
        x360.masker_input(views, names, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        x360.masker_input(views, names, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    px = a.frames * a.views * a.d * a.d
    peak = 6551.4
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    gbs = 15.0 * px / (ms * 1e-3) / 1e9
    print(json.dumps({"op": "views_to_nchw", "metric": "Mpix/s", "value": px / (ms * 1e-3) / 1e6, "ms_per_step": ms,
                      "config": {"frames": a.frames, "views": a.views, "d": a.d, "batch_bytes": 15 * px}, "steps": a.steps,
                      "warmup": a.warmup, "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak}}))


if __name__ == "__main__":
    main()
