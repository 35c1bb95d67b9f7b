#!/usr/bin/env python
"""Small end-to-end invocation for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool racecheck python tools/sanitize_small.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import modulation_mfcc_b200 as mm
from modulation_mfcc_b200 import _lib

sr = 16000
y = mm.synth_batch(0, 6, sr * 3, sr)
for flags in (0, _lib.MMF_FLAG_SCALAR_FFT, _lib.MMF_FLAG_MMA_MEL, _lib.MMF_FLAG_NO_TMA, _lib.MMF_FLAG_UNFUSED_CHANGE):
    res = mm.mfcc_features_batch(y, sr, tStep=0.01, winLen=0.025, n_fft=512, n_mels=40, n_mfcc=13, flags=flags)
    print(flags, float(np.abs(res["totChange"]).sum()), res["modspec"].shape)
# a larger FFT (named barriers per frame group) and a long row (super-block IIR)
y2 = mm.synth_batch(9, 2, 44100 * 2, 44100)
res = mm.mfcc_features_batch(y2, 44100, tStep=0.01, winLen=0.025, n_fft=2048, n_mels=128, n_mfcc=20)
print("2048", float(np.abs(res["totChange"]).sum()))
y3 = mm.synth_batch(11, 1, sr * 90, sr)
res = mm.mfcc_features_batch(y3, sr, tStep=0.01, winLen=0.025, n_fft=512, n_mels=40, n_mfcc=13)
print("long", float(np.abs(res["totChange"]).sum()), res["totChange"].shape)
