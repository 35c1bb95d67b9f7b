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


class PeerGather:
    """All-gather of per-clip feature rows through NVLink peer memory, one push per batch.

    Every rank owns one buffer ``[n_rows, *row_shape]`` allocated as torch symmetric memory, so each rank sees
    every other rank's buffer in its own address space.  ``push(rows, row0)`` copies a finished batch of rows
    into rows ``row0 ..`` of EVERY rank's buffer with plain device-to-device copies on a side stream: the copy
    engines move the data over NVLink / NVSwitch while the SMs run the next batch -- no NCCL kernel competes with
    the persistent compute kernels, and nothing is left to gather at the end but the last batch.
    ``finish()`` waits for this rank's copies and for every peer's (a device-side barrier over the symmetric
    memory signal pads) and returns the local, complete buffer.

    With one process (or when symmetric memory cannot be set up: ``self.mode`` says which) the buffer is a plain
    tensor and ``finish()`` falls back to one in-place ``all_gather_into_tensor``, which needs rank ``r`` to own
    the contiguous row block ``[r*R, (r+1)*R)``, ``R = n_rows / world``.
    """

    def __init__(self, n_rows: int, row_shape, dtype, device, *, group=None, force_collective: bool = False):
        import torch
        import torch.distributed as dist

        self.torch = torch
        self.device = torch.device(device)
        self.group = group
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        shape = (int(n_rows),) + tuple(int(s) for s in row_shape)
        self.mode = "local"
        self.peers = None
        self.hdl = None
        self._pending = []  # (row0, n) pushed locally, for the collective fallback
        if self.world > 1 and not force_collective:
            try:
                import torch.distributed._symmetric_memory as symm_mem

                self.buf = symm_mem.empty(shape, dtype=dtype, device=self.device)
                self.hdl = symm_mem.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
                self.peers = [self.hdl.get_buffer(r, shape, dtype) for r in range(self.world)]
                self.mode = "peer"
            except Exception as e:  # no peer access / unsupported allocator: use the collective
                self.mode = f"collective ({type(e).__name__}: {e})"[:200]
                self.peers = None
        elif self.world > 1:
            self.mode = "collective (forced)"
        if self.peers is None:
            self.buf = torch.empty(shape, dtype=dtype, device=self.device)
        # one side stream per destination: the pushes of a batch run on as many copy engines as the device offers
        # instead of queueing behind each other on one (measured at 8 GPUs: seven 8 MB peer copies per 0.9 ms step
        # on a single stream made the step copy-bound, 1.08 ms)
        self.sides = ([torch.cuda.Stream(device=self.device) for _ in range(max(1, self.world))]
                      if self.device.type == "cuda" else None)

    def push(self, rows, row0: int):
        """Queue ``rows`` (a tensor on this rank's device, produced on the current stream) for rows
        ``row0 : row0 + len(rows)`` of every rank's buffer.  Returns immediately."""
        torch = self.torch
        n = int(rows.shape[0])
        if self.peers is None:
            self.buf[row0 : row0 + n].copy_(rows, non_blocking=True)
            self._pending.append((int(row0), n))
            return
        cur = torch.cuda.current_stream(self.device)
        for i in range(self.world):
            r = (self.rank + i) % self.world  # own copy first, then the peers round-robin
            side = self.sides[i]
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                self.peers[r][row0 : row0 + n].copy_(rows, non_blocking=True)
            rows.record_stream(side)

    def finish(self):
        """Block the current stream until every rank's rows have landed in this rank's buffer; returns it."""
        torch = self.torch
        if self.peers is not None:
            cur = torch.cuda.current_stream(self.device)
            for side in self.sides:
                cur.wait_stream(side)
            self.hdl.barrier()  # device-side: all ranks' copies are complete and visible
            return self.buf
        if self.world > 1:
            import torch.distributed as dist

            # collective fallback: rank r must own the contiguous block [r*R, (r+1)*R), R = n_rows / world
            R = self.buf.shape[0] // self.world
            if R * self.world != self.buf.shape[0] or any(a < self.rank * R or a + n > (self.rank + 1) * R for a, n in self._pending):
                raise RuntimeError("collective fallback of PeerGather needs equal contiguous row blocks per rank")
            dist.all_gather_into_tensor(self.buf, self.buf[self.rank * R : (self.rank + 1) * R], group=self.group)
        self._pending.clear()
        return self.buf
