"""N > 1 on real GPUs: two ranks (one per GPU) over NCCL.  The frame-sharded run must produce the same views and the same
keep table as a 1-GPU run of the same job (the one collective of the design, the all-gather of the Laplacian sum pairs,
runs on NVLink here; tests/test_sharding_gloo.py covers the host logic on CPU), and so must the view-sharded run of a
single-frame job.  Skipped on boxes with fewer than two GPUs."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.skipif(_gpus() < 2, reason="needs two GPUs (NCCL refuses two ranks on one device)")
def test_two_ranks_over_nccl_match_one_gpu(x360, tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _nccl_worker import job
    settings, frames = job()
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "_nccl_worker.py"), str(tmp_path)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]

    # the same job on one GPU, in this process
    with x360.ExtractionSession(settings, 256, 512, device=0, max_batch=2) as sess:
        want_views, want_scores = sess.extract_batch(frames)
        want_keep = np.array(sess.blur.replay(want_scores), dtype=bool)
        one_views, one_scores = sess.extract_batch(frames[:1])
    with x360.ExtractionSession(settings, 256, 512, device=0) as sess1:
        one_keep = np.array(sess1.blur.replay(one_scores), dtype=bool)
    assert want_keep.any() and not want_keep.all(), "the job must exercise accept AND reject"

    ranks = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    assert all(str(z["backend"]) == "nccl" for z in ranks)
    got = np.concatenate([z["views"] for z in ranks], axis=0)
    assert [int(z["lo"]) for z in ranks] == [x360.shard_bounds(len(frames), world, r)[0] for r in range(world)]
    assert np.array_equal(got, want_views), "frame-sharded views differ from the 1-GPU run"
    for r, z in enumerate(ranks):
        assert np.array_equal(z["keep"], want_keep), f"rank {r}: keep table differs from the 1-GPU replay"
        s = np.asarray(z["gs"], dtype=np.float64) / (160 * 160)
        assert np.allclose(np.asarray(z["gs2"], dtype=np.float64) / (160 * 160) - s * s, want_scores, rtol=1e-12, atol=1e-9), f"rank {r}: gathered sums"
    # single-frame job: sharded by view
    assert all(str(z["axis"]) == "view" for z in ranks)
    got1 = np.concatenate([z["v_views"] for z in ranks], axis=1)
    assert np.array_equal(got1, one_views), "view-sharded views differ from the 1-GPU run"
    for r, z in enumerate(ranks):
        assert np.array_equal(z["v_keep"], one_keep), f"rank {r}: view-sharded keep table"
