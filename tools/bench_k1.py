#!/usr/bin/env python
"""Time the fused STFT/mel kernel alone (plan.logmel) for a few shapes and flag sets.

    python tools/bench_k1.py [--clips 1024] [--flags 0,16] [--shape cfg2|gui|cfg3]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import modulation_mfcc_b200 as mm

SHAPES = {
    # name: (sr, seconds, n_fft, winLen, tStep, n_mels, n_mfcc, fmin, fmax)
    "cfg2": (16000, 10.0, 512, 0.025, 0.01, 40, 13, 0.0, 8000.0),
    "gui": (10000, 10.0, 512, 0.025, 0.005, 128, 13, 100.0, 10000.0),
    "cfg3": (44100, 10.0, 2048, 0.025, 0.01, 128, 20, 0.0, 22050.0),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=1024)
    ap.add_argument("--flags", default="0")
    ap.add_argument("--shape", default="cfg2")
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    sr, secs, n_fft, winLen, tStep, n_mels, n_mfcc, fmin, fmax = SHAPES[a.shape]
    n = int(sr * secs)
    clips = a.clips if a.shape != "cfg3" else min(a.clips, 512)
    pcm = mm.synth_batch_device(clips, n, sr, seed=1, device=torch.device("cuda", 0))
    for fl in [int(x) for x in a.flags.split(",")]:
        win, hop = mm.frame_sizes(sr, winLen, tStep)
        plan = mm.get_plan(mm.MfccConfig(sr, n_fft, win, hop, n_mels, n_mfcc, fmin, fmax, flags=fl))
        for _ in range(3):
            plan.logmel(pcm)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            plan.logmel(pcm)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.iters
        T = plan.num_frames(n)
        gb = clips * (4 * n + 4 * n_mels * T) / 1e9
        print(f"{a.shape} clips {clips} flags {fl}: {ms:.4f} ms  {gb / (ms * 1e-3):.0f} GB/s algorithmic")


if __name__ == "__main__":
    main()
