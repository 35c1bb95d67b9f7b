#!/usr/bin/env python
"""Time mmf_stft_power (transform only, power spectrum to HBM) for plan flag sets, e.g. the FP32 register
FFT (0) against the tcgen05 transform (MMF_FLAG_TC_FFT = 256).

    python tools/bench_power.py [--clips 1024] [--flags 0,256]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import modulation_mfcc_b200 as mm


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=1024)
    ap.add_argument("--flags", default="0,256")
    ap.add_argument("--iters", type=int, default=10)
    a = ap.parse_args()
    sr, n = 16000, 160000
    pcm = mm.synth_batch_device(a.clips, n, sr, seed=1, device=torch.device("cuda", 0))
    win, hop = mm.frame_sizes(sr, 0.025, 0.01)
    for fl in [int(x) for x in a.flags.split(",")]:
        plan = mm.get_plan(mm.MfccConfig(sr, 512, win, hop, 40, 13, 0.0, 8000.0, flags=fl))
        for _ in range(2):
            out = plan.stft_power(pcm)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            out = plan.stft_power(pcm)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.iters
        print(f"flags {fl}: {ms:.4f} ms per {a.clips} clips (power write {out.numel() * 4 / 1e9:.2f} GB)")
        del out


if __name__ == "__main__":
    main()
