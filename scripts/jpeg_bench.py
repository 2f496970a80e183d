"""Device-resident timing of the JPEG encoder (x360_jpeg_encode_device) by batch size: 1600 x 1600 views of the bench clip
(SURVEY 8(d) smooth distribution) at quality 95, CUDA events on the launch stream, 10 timed encodes after 3 warm-ups.

    python scripts/jpeg_bench.py [--sizes 8,16,32,64,256] [--once N]     (--once: two encodes of N views and exit, for ncu)
"""
import argparse
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def batch(n, d):
    base = bench.base_frame(0, d, d, "smooth")
    imgs = torch.from_numpy(np.stack([np.roll(base, 37 * i, axis=1) for i in range(min(n, 8))])).cuda()
    return imgs.repeat((n + imgs.shape[0] - 1) // imgs.shape[0], 1, 1, 1)[:n].contiguous()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="8,16,32,64,256")
    ap.add_argument("--once", type=int, default=0)
    ap.add_argument("--d", type=int, default=1600)
    args = ap.parse_args()
    x = importlib.import_module("360extractor_b200")
    d = args.d
    sizes = [args.once] if args.once else [int(s) for s in args.sizes.split(",")]
    for n in sizes:
        imgs = batch(n, d)
        enc = x.JpegEncoder(d, d, 95, max_images=n)
        out = torch.empty(enc.max_bytes(n), dtype=torch.uint8, device="cuda")
        off = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
        for _ in range(2 if args.once else 3):
            enc.encode_device(imgs, out, off)
        torch.cuda.synchronize()
        if args.once:
            return
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            enc.encode_device(imgs, out, off)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(json.dumps({"views": n, "size": d, "quality": 95, "ms": round(ms, 4), "gpix_per_s": round(n * d * d / ms / 1e6, 2),
                          "bytes_per_pixel": round(int(off[n].item()) / (n * d * d), 4)}))
        enc.close()


if __name__ == "__main__":
    main()
