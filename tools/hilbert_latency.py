#!/usr/bin/env python
"""Latency of the Hilbert envelope (calc.py:284-286) for one clip, after warm-up: host call and GPU timeline."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import modulation_mfcc_b200 as mm

for n, sr in ((100_000, 10_000), (160_000, 16_000), (131_072, 16_000)):
    y = mm.synth_clip(1, n, sr)
    for _ in range(10):
        mm.calculate_amplitude_envelope(y, sr, method="Hilb")
    ts = []
    for _ in range(30):
        t0 = time.perf_counter()
        mm.calculate_amplitude_envelope(y, sr, method="Hilb")
        ts.append(time.perf_counter() - t0)
    print(f"n = {n}: median {1e3 * np.median(ts):.3f} ms, min {1e3 * min(ts):.3f} ms, max {1e3 * max(ts):.3f} ms")
