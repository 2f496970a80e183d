"""Rank program of tests/test_multi_gpu_nccl.py (launched with torch.distributed.run, one rank per GPU, NCCL).

Every rank builds the same ExtractionSession, runs the frame-sharded and the view-sharded path on the same seeded
frames and writes what it got to <out_dir>/rank<r>.npz; the pytest process compares that with its own 1-GPU run."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def job():
    settings = {"resolution": 160, "fov": 90, "camera_count": 6, "pitch_offset": 0, "layout_mode": "cube",
                "interpolation_mode": "linear", "blur_filter_enabled": True, "smart_blur_enabled": True, "blur_threshold": 9000.0}
    rng = np.random.default_rng(77)
    frames = rng.integers(0, 256, (7, 256, 512, 3), dtype=np.uint8)
    frames[2:5] = (frames[2:5].astype(np.uint16) * 5 // 16 + 90).astype(np.uint8)      # three low-contrast frames: rejected, then force-accepted
    return settings, frames


def main():
    out_dir = sys.argv[1]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = importlib.import_module("360extractor_b200")
    settings, frames = job()
    with pkg.ExtractionSession(settings, 256, 512, device=local, max_batch=2) as sess:
        lo, views, keep = sess.process_sharded(frames, rank, world)
        axis, v_lo, v_views, v_keep = sess.process_distributed(frames[:1], rank, world)
        # the device-resident form the bench uses: CUDA int64 sums straight into the NCCL all-gather
        f_lo, f_hi = pkg.shard_bounds(len(frames), world, rank)
        d_fr = torch.from_numpy(frames[f_lo:f_hi]).cuda()
        d_out = torch.empty(sess.extractor.views_shape(f_hi - f_lo), dtype=torch.uint8, device="cuda")
        d_s = torch.zeros((f_hi - f_lo, sess.extractor.n_views), dtype=torch.int64, device="cuda")
        d_s2 = torch.zeros_like(d_s)
        sess.extractor.extract_device(d_fr, d_out, d_s, d_s2, stream=torch.cuda.current_stream())
        gs, gs2 = pkg.gather_sums(d_s, d_s2, len(frames))
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), lo=lo, views=views, keep=keep, axis=axis, v_lo=v_lo, v_views=v_views, v_keep=v_keep,
             gs=gs, gs2=gs2, backend=dist.get_backend())
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
