"""N > 1 host logic on CPU: frame sharding, the score all-gather (gloo, world_size 2 and 3) and the
blur accept/reject replay that every rank runs over the gathered table.  No CUDA involved: ranks
feed synthetic int64 Laplacian sums, exactly what the kernel hands to ``gather_sums`` on a GPU box."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_and_balance(x360):
    for n in (0, 1, 5, 7, 32, 256, 257):
        for world in (1, 2, 3, 4, 8):
            bounds = [x360.shard_bounds(n, world, r) for r in range(world)]
            assert bounds[0][0] == 0 and bounds[-1][1] == n
            assert all(bounds[i][1] == bounds[i + 1][0] for i in range(world - 1))       # contiguous blocks
            sizes = [hi - lo for lo, hi in bounds]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
            assert sizes == x360.shard_sizes(n, world)
    with pytest.raises(ValueError):
        x360.shard_bounds(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _synthetic_sums(F, V, npix):
    """Deterministic per-(frame, view) sums with a few blurry stretches."""
    rng = np.random.default_rng(11)
    var = rng.uniform(150, 900, (F, V))
    if F >= 9:
        var[5:9] = rng.uniform(5, 60, (4, V))             # a run of blurry frames -> forced accepts in smart mode
    s = rng.integers(-50, 50, (F, V)).astype(np.int64) * 1000
    s2 = np.rint((var + (s / npix) ** 2) * npix).astype(np.int64)
    return s, s2


def _worker(rank, world, port, F, V, npix, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = importlib.import_module("360extractor_b200")
    s, s2 = _synthetic_sums(F, V, npix)
    lo, hi = pkg.shard_bounds(F, world, rank)
    gs, gs2 = pkg.gather_sums(torch.from_numpy(s[lo:hi].copy()), torch.from_numpy(s2[lo:hi].copy()), F)
    scores = pkg.scores_from_sums(gs, gs2, npix)
    keep = np.array(pkg.BlurFilter(True, True, 100.0).replay(scores), dtype=bool)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), gs=gs, gs2=gs2, keep=keep)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_score_gather_and_replay_agree_on_every_rank(x360, tmp_path, world):
    F, V, npix = 13, 6, 64 * 64                            # 13 frames over 2 / 3 ranks: ragged shards
    mp.spawn(_worker, args=(world, _free_port(), F, V, npix, str(tmp_path)), nprocs=world, join=True)
    s, s2 = _synthetic_sums(F, V, npix)
    want_scores = x360.scores_from_sums(s, s2, npix)
    want_keep = np.array(x360.BlurFilter(True, True, 100.0).replay(want_scores), dtype=bool)
    assert not want_keep.all() and want_keep.any()
    for r in range(world):
        z = np.load(tmp_path / f"rank{r}.npz")
        assert np.array_equal(z["gs"], s) and np.array_equal(z["gs2"], s2), f"rank {r} gathered table"
        assert np.array_equal(z["keep"], want_keep), f"rank {r} keep table"


def test_gather_without_process_group_is_identity(x360):
    s, s2 = _synthetic_sums(4, 3, 100)
    gs, gs2 = x360.gather_sums(s, s2, 4)
    assert np.array_equal(gs, s) and np.array_equal(gs2, s2)


def _view_worker(rank, world, port, F, V, npix, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = importlib.import_module("360extractor_b200")
    s, s2 = _synthetic_sums(F, V, npix)
    lo, hi = pkg.shard_bounds(V, world, rank)                     # this rank's block of the VIEW list
    gs, gs2 = pkg.gather_sums_by_view(s[:, lo:hi].copy(), s2[:, lo:hi].copy(), V)
    keep = np.array(pkg.BlurFilter(True, True, 100.0).replay(pkg.scores_from_sums(gs, gs2, npix)), dtype=bool)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), gs=gs, gs2=gs2, keep=keep)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,F,V", [(2, 1, 5), (3, 2, 4), (3, 1, 2)])
def test_view_sharded_gather_for_single_frame_jobs(x360, tmp_path, world, F, V):
    """Fewer frames than ranks (a single 360 image): views are sharded instead; ragged blocks, and with
    V < world a rank without any view still takes part in the collective."""
    assert x360.shard_axis(F, world) == "view" and x360.shard_axis(world, world) == "frame"
    npix = 48 * 48
    mp.spawn(_view_worker, args=(world, _free_port(), F, V, npix, str(tmp_path)), nprocs=world, join=True)
    s, s2 = _synthetic_sums(F, V, npix)
    want_keep = np.array(x360.BlurFilter(True, True, 100.0).replay(x360.scores_from_sums(s, s2, npix)), dtype=bool)
    for r in range(world):
        z = np.load(tmp_path / f"rank{r}.npz")
        assert np.array_equal(z["gs"], s) and np.array_equal(z["gs2"], s2), f"rank {r} gathered table"
        assert np.array_equal(z["keep"], want_keep), f"rank {r} keep table"
