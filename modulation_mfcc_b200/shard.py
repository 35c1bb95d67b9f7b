"""Corpus sharding across the GPUs of one box.

Clips are independent units: rank ``r`` of ``world`` processes owns the contiguous
block ``clips[lo:hi]`` and runs the whole path on it with no data-path
collective.  The only collective is the final gather of per-clip features
(``torch.distributed.all_gather`` — NCCL on the GPU box, gloo in the CPU tests).
"""

from __future__ import annotations

import numpy as np


def shard_range(n_clips: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block partition; the first ``n_clips % world`` ranks own one extra clip."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    if n_clips < 0:
        raise ValueError("n_clips must be non-negative")
    base, extra = divmod(n_clips, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n_clips: int, world: int) -> list[int]:
    return [hi - lo for lo, hi in (shard_range(n_clips, r, world) for r in range(world))]


def gather_features(local, n_clips: int, *, group=None):
    """All-gather the per-clip feature rows of every rank into ``[n_clips, ...]``.

    ``local`` is this rank's ``[hi - lo, ...]`` tensor (any device).  Shards may be
    ragged (``n_clips % world != 0``): rows are padded to the largest shard for the
    collective and trimmed afterwards, so clip order equals the unsharded order.
    """
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        if local.shape[0] != n_clips:
            raise ValueError("single process must hold every clip")
        return local
    world = dist.get_world_size(group)
    sizes = shard_sizes(n_clips, world)
    rank = dist.get_rank(group)
    if local.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank} holds {local.shape[0]} clips, its shard is {sizes[rank]}")
    width = max(sizes)
    if local.shape[0] < width:
        pad = torch.zeros((width - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        local = torch.cat([local, pad], dim=0)
    local = local.contiguous()
    out = torch.empty((world * width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local, group=group)
    if all(s == width for s in sizes):
        return out
    keep = np.concatenate([np.arange(r * width, r * width + s) for r, s in enumerate(sizes)])
    return out[torch.as_tensor(keep, device=out.device)]
