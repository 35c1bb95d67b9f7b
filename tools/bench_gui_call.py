#!/usr/bin/env python
"""Latency of the drop-in call the GUI makes (script/main.py:750-769): one 10 s clip at 10 kHz,
host numpy array in, host numpy arrays out."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import modulation_mfcc_b200 as mm

KW = dict(channelN=0, tStep=0.005, winLen=0.025, n_mfcc=13, n_fft=512, minFreq=100, maxFreq=10000, removeFirst=1,
          filtCutoff=12, filtOrd=6, diffMethod="grad", outFilter="iir", outFiltType="low", outFiltCutOff=[12],
          outFiltLen=6, outFiltPolyOrd=3)
y = mm.synth_clip(1, 100_000, 10_000)
for _ in range(5):
    mm.get_MFCCS_change(y, 10_000, **KW)
ts = []
for _ in range(50):
    t0 = time.perf_counter()
    tot, T = mm.get_MFCCS_change(y, 10_000, **KW)
    ts.append(time.perf_counter() - t0)
print(f"get_MFCCS_change, one 10 s clip @ 10 kHz (GUI defaults): median {1e3 * np.median(ts):.3f} ms, "
      f"min {1e3 * min(ts):.3f} ms, T = {len(T)} frames")
a, t = mm.calculate_amplitude_envelope(y, 10_000, method="RMS")
ts = []
for _ in range(20):
    t0 = time.perf_counter()
    mm.calculate_amplitude_envelope(y, 10_000, method="RMS")
    ts.append(time.perf_counter() - t0)
print(f"calculate_amplitude_envelope RMS: median {1e3 * np.median(ts):.3f} ms")
ts = []
for _ in range(5):
    t0 = time.perf_counter()
    mm.calculate_amplitude_envelope(y, 10_000, method="Hilb")
    ts.append(time.perf_counter() - t0)
print(f"calculate_amplitude_envelope Hilb: median {1e3 * np.median(ts):.3f} ms")
