"""Deterministic synthetic audio: white noise + FM tone with 4 Hz AM.

Clip ``i`` is ``0.1*N(0,1) + 0.3*sin(2*pi*(200 + 50*sin(2*pi*3t))*t) * (0.6 + 0.4*sin(2*pi*4t))``
clipped to [-1, 1], ``rng = np.random.default_rng(1234 + i)``, float32 mono.
The 4 Hz amplitude modulation is speech-rate, so the MFCC modulation spectrum
is non-trivial.  Host (numpy) generator for parity subsets and CPU baselines;
``synth_batch_device`` is the same recipe on the GPU for corpus-scale runs.
"""

from __future__ import annotations

import numpy as np


def synth_clip(i: int, n_samples: int, sr: float) -> np.ndarray:
    rng = np.random.default_rng(1234 + int(i))
    t = np.arange(n_samples, dtype=np.float64) / sr
    noise = 0.1 * rng.standard_normal(n_samples)
    fm = 200.0 + 50.0 * np.sin(2 * np.pi * 3.0 * t)
    tone = 0.3 * np.sin(2 * np.pi * fm * t) * (0.6 + 0.4 * np.sin(2 * np.pi * 4.0 * t))
    return np.clip(noise + tone, -1.0, 1.0).astype(np.float32)


def synth_batch(first: int, n_clips: int, n_samples: int, sr: float) -> np.ndarray:
    return np.stack([synth_clip(first + i, n_samples, sr) for i in range(n_clips)])


def synth_batch_device(n_clips: int, n_samples: int, sr: float, *, seed: int, device):
    """Same recipe generated on ``device`` with torch (different noise stream than
    numpy's; used for corpus-scale throughput runs where host generation plus
    H2D would dominate).  Returns a float32 ``[n_clips, n_samples]`` tensor."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    t = torch.arange(n_samples, device=device, dtype=torch.float64) / sr
    fm = 200.0 + 50.0 * torch.sin(2 * torch.pi * 3.0 * t)
    tone = (0.3 * torch.sin(2 * torch.pi * fm * t) * (0.6 + 0.4 * torch.sin(2 * torch.pi * 4.0 * t))).to(torch.float32)
    out = torch.empty((n_clips, n_samples), device=device, dtype=torch.float32)
    step = 64
    for s in range(0, n_clips, step):
        e = min(n_clips, s + step)
        noise = torch.randn((e - s, n_samples), device=device, dtype=torch.float32, generator=g)
        out[s:e] = torch.clamp(0.1 * noise + tone[None, :], -1.0, 1.0)
    return out
