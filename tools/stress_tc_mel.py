#!/usr/bin/env python
"""Randomised soak of the tcgen05 mel kernel: many (n_clips, n_samples) shapes -- single-tile blocks, grids smaller
than the SM count, ragged ends -- against the CUDA-core kernel (1e-4 on mel power) and against itself (bit-identical on
a second launch).  A hang shows up as the surrounding `timeout`."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import modulation_mfcc_b200 as mm
from modulation_mfcc_b200 import _lib

rng = np.random.default_rng(2026)
sr = 16000
cfgs = [
    mm.MfccConfig(sr, 512, 400, 160, 40, 13, 0.0, 8000.0),
    mm.MfccConfig(sr, 512, 512, 128, 64, 13, 50.0, 7000.0),
    mm.MfccConfig(10000, 512, 250, 50, 128, 13, 100.0, 10000.0),
    mm.MfccConfig(sr, 512, 320, 161, 26, 13, 0.0, 8000.0),
]
t_end = time.time() + float(sys.argv[1]) if len(sys.argv) > 1 else time.time() + 40.0
n = 0
worst = 0.0
while time.time() < t_end:
    cfg = cfgs[n % len(cfgs)]
    n_clips = int(rng.choice([1, 2, 3, 7, 33, 148, 149, 300]))
    n_samples = int(rng.choice([1, 159, 5000, 10239, 10240, 20481, 48000, 160000, 163841]))
    if n_clips * n_samples > 40_000_000:
        n_clips = max(1, 40_000_000 // n_samples)
    y = torch.randn(n_clips, n_samples, device="cuda") * float(rng.choice([1e-3, 0.1, 30.0]))
    a, ka = mm.get_plan(cfg).logmel(y)
    a2, ka2 = mm.get_plan(cfg).logmel(y)
    assert torch.equal(a, a2) and torch.equal(ka, ka2), ("not deterministic", cfg, n_clips, n_samples)
    b, kb = mm.get_plan(mm.plan.replace(cfg, flags=_lib.MMF_FLAG_NO_TC_MEL)).logmel(y)
    ok = b > -99.0
    d = (a - b).abs()[ok]
    err = float((10.0 ** (d / 10.0) - 1.0).max()) if d.numel() else 0.0
    worst = max(worst, err)
    assert err <= 1e-4, (err, cfg, n_clips, n_samples)
    assert torch.isfinite(a).all()
    n += 1
print(f"{n} shapes ok, worst mel-power deviation from the CUDA-core kernel {worst:.2e}")
