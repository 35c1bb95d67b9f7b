#!/usr/bin/env python
"""Error of the modulation-spectrum kernels against the float64 oracle on identical float32 input:
tcgen05 GEMM (flags 0, default) vs FP32 register FFT (MMF_FLAG_NO_TC_MODSPEC = 512)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

import modulation_mfcc_b200 as mm
from oracle import mfcc_oracle as oracle

rng = np.random.default_rng(11)
M = (rng.standard_normal((4, 13, 1001)).cumsum(axis=-1) * 0.5).astype(np.float32)
for win_s, hop_s, fr in [(1.0, 0.5, 100.0), (1.0, 0.01, 100.0), (0.5, 0.25, 100.0)]:
    Lw, Hw, nfft, n_win = mm.modspec_sizes(1001, fr, win_s, hop_s)
    bins = mm.band_bins(nfft, fr)
    for fl in (0, 512):
        plan = mm.get_plan(mm.MfccConfig(16000, 512, 400, 160, 40, 13, 0.0, 8000.0, flags=fl))
        mag, band = plan.modspec(torch.as_tensor(M).cuda(), Lw, Hw, nfft, bins)
        mag, band = mag.cpu().numpy(), band.cpu().numpy()
        em, eb, sb = 0.0, 0.0, 0.0
        for i in range(4):
            rm, rb, _ = oracle.modulation_spectrum(M[i], fr, mod_win_s=win_s, mod_hop_s=hop_s)
            em = max(em, float(np.max(np.abs(mag[i] - rm))))
            rel = (band[i].astype(np.float64) - rb) / np.maximum(rb, 1e-30)
            eb = max(eb, float(np.max(np.abs(rel))))
            sb += float(np.mean(rel)) / 4
        print(f"win {Lw} hop {Hw} nfft {nfft} flags {fl}: mag max abs err {em:.3e} (peak {np.abs(mag).max():.1f}), "
              f"band max rel err {eb:.3e}, mean rel err {sb:+.3e}")
