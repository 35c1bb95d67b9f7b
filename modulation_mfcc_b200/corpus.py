"""Corpus-scale driver: the whole path over a device-resident shard of clips, batch after batch.

BASELINE configs[4] (a 100k-clip corpus over the GPUs of one box): rank ``r`` owns the
contiguous block ``shard_range(n_clips, r, world)`` of clip indices and runs the five kernels of the path on it
in batches.  Nothing synchronises with the host between batches, every output of the shard is preallocated, the
plan's workspace is reused, and each finished batch of per-clip curves is handed to a :class:`PeerGather`, whose
copy engines spread it to the peers over NVLink while the next batch computes.
"""

from __future__ import annotations

from .api import FeatureExtractor
from .shard import PeerGather, shard_range


class CorpusRunner:
    def __init__(self, fx: FeatureExtractor, n_clips: int, n_samples: int, *, batch: int = 1024, rank: int = 0,
                 world: int = 1, gather: PeerGather | None = None, want_band_energy: bool = True):
        import torch

        self.fx, self.batch = fx, int(batch)
        self.n_clips, self.n_samples = int(n_clips), int(n_samples)
        self.lo, self.hi = shard_range(self.n_clips, rank, world)
        self.T = fx.plan.num_frames(self.n_samples)
        dev = fx.plan.device
        self.gather = gather
        n = self.hi - self.lo
        # per-clip outputs of the shard; the curves of ALL clips live in the gather buffer
        self.tot = None if gather is not None else torch.empty((n, self.T), device=dev, dtype=torch.float64)
        self.band = None
        self.mod = None
        if want_band_energy:
            Lw, Hw, nfft, bins = fx.modspec_geometry(self.T)
            self.mod = (Lw, Hw, nfft, bins)
            n_win = 1 + (self.T - Lw) // Hw if self.T >= Lw else 0
            self.band = torch.empty((n, n_win, len(bins)), device=dev, dtype=torch.float32)

    def run(self, pcm_shard):
        """``pcm_shard``: float32 CUDA tensor ``[hi - lo, n_samples]``.  Queues every batch on the current
        stream and returns without synchronising; call ``finish()`` for the gathered curves."""
        plan, prm = self.fx.plan, self.fx.prm
        n = self.hi - self.lo
        if tuple(pcm_shard.shape) != (n, self.n_samples):
            raise ValueError(f"shard must be [{n}, {self.n_samples}]")
        for b0 in range(0, n, self.batch):
            nb = min(self.batch, n - b0)
            res = plan.mfcc_change(pcm_shard[b0 : b0 + nb], prm, want_mfcc=self.mod is not None)
            if self.mod is not None:
                Lw, Hw, nfft, bins = self.mod
                _, band = plan.modspec(res["mfcc"], Lw, Hw, nfft, bins, want_mag=False)
                self.band[b0 : b0 + nb].copy_(band, non_blocking=True)
            if self.gather is not None:
                self.gather.push(res["totChange"], self.lo + b0)
            else:
                self.tot[b0 : b0 + nb].copy_(res["totChange"], non_blocking=True)

    def finish(self):
        """Curves of all ``n_clips`` clips (``[n_clips, T]`` float64) when a gather is attached, else the shard's."""
        return self.gather.finish() if self.gather is not None else self.tot
