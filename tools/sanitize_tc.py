#!/usr/bin/env python
"""Small invocation of the round-2 kernels for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool racecheck python tools/sanitize_tc.py

tcgen05 mel kernel (TMA and plain loader, ragged lengths, one- and two-tile blocks), packed clamp + DCT kernel,
per-clip kernel + batched output filter (>= 32 clips)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import modulation_mfcc_b200 as mm
from modulation_mfcc_b200 import _lib

sr = 16000
for n, clips in ((sr * 2 + 37, 3), (sr // 2, 34), (9000, 2)):
    y = mm.synth_batch(0, clips, n, sr)
    for flags in (0, _lib.MMF_FLAG_NO_TMA, _lib.MMF_FLAG_NO_TC_MEL):
        res = mm.mfcc_features_batch(y, sr, tStep=0.01, winLen=0.025, n_fft=512, n_mels=40, n_mfcc=13, flags=flags)
        print(n, clips, flags, float(np.abs(res["totChange"]).sum()), float(np.abs(res["mfcc"]).sum()))
