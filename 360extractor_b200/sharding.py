"""Sharding across the GPUs of one box: by frame, or — for jobs with fewer frames than GPUs — by view.

Every (frame, view, tile) is independent, so the work is split into contiguous index blocks, one
per rank, with no pixel traffic between GPUs: frames of the sampled frame list (after the
``frame_idx % interval`` gate, src/core/processor.py:286), or the views of the per-frame loop
(src/core/processor.py:327-330) when a job is a single image (every rank then uploads the same
frame; nothing is broadcast).  The only exchange is the per-view blur score: the accept/reject
machine of src/core/processor.py:345-367 is sequential over (frame, view), so the int64 sum pairs
are all-gathered (a few KB, NCCL on GPUs / gloo in the CPU tests) and every rank replays the
machine over the full, ordered score table.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n_items: int, world_size: int, rank: int):
    """Contiguous block ``[lo, hi)`` of ``n_items`` for ``rank``; the first ``n % world`` ranks get one more."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank {rank} / world {world_size}")
    base, extra = divmod(int(n_items), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n_items: int, world_size: int):
    return [shard_bounds(n_items, world_size, r)[1] - shard_bounds(n_items, world_size, r)[0] for r in range(world_size)]


def shard_axis(n_frames: int, world_size: int) -> str:
    """'frame' when every rank can get at least one frame, else 'view' (single-image jobs)."""
    return "frame" if n_frames >= world_size else "view"


def gather_sums_by_view(sum_lap, sum_lap2, n_views_total: int, group=None):
    """View-sharded counterpart of ``gather_sums``: every rank holds ``[F][V_local]`` for its view block
    (``shard_bounds(V, world, rank)``); returns the full ``[F][V]`` tables on every rank."""
    import torch
    t = (lambda a: a.transpose(0, 1).contiguous()) if torch.is_tensor(sum_lap) else (lambda a: np.ascontiguousarray(np.asarray(a).T))
    s, s2 = gather_sums(t(sum_lap), t(sum_lap2), n_views_total, group=group)
    return np.ascontiguousarray(s.T), np.ascontiguousarray(s2.T)


def gather_sums(sum_lap, sum_lap2, n_frames_total: int, group=None):
    """All-gather the per-rank int64 ``[F_local][V]`` sum pairs into the global ``[F][V]`` tables.

    Inputs may be numpy arrays or torch tensors (CPU for gloo, CUDA for nccl).  Ranks may hold
    different frame counts (``shard_bounds``); rows are padded to the largest shard for the
    collective and trimmed afterwards.  Returns numpy int64 arrays on every rank."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return np.asarray(_to_numpy(sum_lap)), np.asarray(_to_numpy(sum_lap2))
    world = dist.get_world_size(group)
    sizes = shard_sizes(n_frames_total, world)
    local = torch.as_tensor(sum_lap) if not torch.is_tensor(sum_lap) else sum_lap
    local2 = torch.as_tensor(sum_lap2) if not torch.is_tensor(sum_lap2) else sum_lap2
    device = local.device
    if dist.get_backend(group) == "nccl" and device.type != "cuda":
        device = torch.device("cuda", torch.cuda.current_device())
    V = int(local.shape[1]) if local.ndim == 2 else 0
    pad = max(sizes)
    send = torch.zeros((2, pad, V), dtype=torch.int64, device=device)
    n_local = int(local.shape[0])
    send[0, :n_local] = local.to(device)
    send[1, :n_local] = local2.to(device)
    recv = torch.empty((world, 2, pad, V), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(recv.view(-1), send.view(-1), group=group) if dist.get_backend(group) == "nccl" \
        else dist.all_gather(list(recv.unbind(0)), send, group=group)
    recv = recv.cpu().numpy()
    s = np.concatenate([recv[r, 0, :sizes[r]] for r in range(world)], axis=0)
    s2 = np.concatenate([recv[r, 1, :sizes[r]] for r in range(world)], axis=0)
    return s, s2


def _to_numpy(a):
    try:
        import torch
        if torch.is_tensor(a):
            return a.detach().cpu().numpy()
    except ImportError:  # pragma: no cover
        pass
    return np.asarray(a)


__all__ = ["shard_bounds", "shard_sizes", "shard_axis", "gather_sums", "gather_sums_by_view"]
