"""Host-side logic of the multi-GPU path: clip partition + final gather.

The GPU box runs the gather over NCCL; here the same code runs with two CPU
processes over gloo (world_size 2), which is what SURVEY.md §8e asks the CPU
suite to cover.
"""

import os
import socket

import numpy as np
import pytest

from modulation_mfcc_b200.shard import gather_features, shard_range, shard_sizes


@pytest.mark.parametrize("n,world", [(0, 1), (1, 1), (5, 2), (1024, 8), (100_000, 8), (7, 8), (100_000, 3)])
def test_shard_range_is_a_contiguous_partition(n, world):
    ranges = [shard_range(n, r, world) for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == n
    for (a0, a1), (b0, b1) in zip(ranges, ranges[1:]):
        assert a1 == b0 and a0 <= a1
    sizes = shard_sizes(n, world)
    assert sum(sizes) == n and max(sizes) - min(sizes) <= 1


def test_shard_range_rejects_bad_rank():
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)
    with pytest.raises(ValueError):
        shard_range(-1, 0, 1)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _per_clip_rows(lo, hi, width):
    # stand-in for the per-clip feature (totChange row): depends only on the clip index
    idx = np.arange(lo, hi, dtype=np.float64)[:, None]
    return idx * 1000.0 + np.arange(width, dtype=np.float64)[None, :]


def _worker(rank, world, port, n_clips, width, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_range(n_clips, rank, world)
        local = torch.from_numpy(_per_clip_rows(lo, hi, width))
        full = gather_features(local, n_clips)
        q.put((rank, full.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_clips", [8, 5])
def test_gather_world2_gloo(n_clips):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port, width, world = _free_port(), 7, 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_clips, width, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = _per_clip_rows(0, n_clips, width)
    for r in range(world):
        np.testing.assert_array_equal(got[r], want)


def test_gather_single_process_passthrough():
    import torch

    x = torch.arange(12.0).reshape(4, 3)
    assert gather_features(x, 4) is x
    with pytest.raises(ValueError):
        gather_features(x, 5)
