#!/usr/bin/env python
"""Time plan.mfcc_change (the whole get_MFCCS_change on device PCM) with and without the
MFCC / delta copies in HBM, for the per-clip kernel with the DCT folded in (flags 0) and with
the separate MFCC kernel (MMF_FLAG_SEPARATE_MFCC = 64).

    python tools/bench_change_only.py [--clips 1024] [--flags 0,64]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scipy.signal
import torch

import modulation_mfcc_b200 as mm


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=1024)
    ap.add_argument("--flags", default="0,64")
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    sr, n = 16000, 160000
    pcm = mm.synth_batch_device(a.clips, n, sr, seed=1, device=torch.device("cuda", 0))
    win, hop = mm.frame_sizes(sr, 0.025, 0.01)
    sos = scipy.signal.butter(6, 12.0, "low", fs=sr / hop, output="sos")
    prm = mm.plan.make_change_params(sos, remove_first=1, diff_method=0, out_sos=sos)
    for fl in [int(x) for x in a.flags.split(",")]:
        plan = mm.get_plan(mm.MfccConfig(sr, 512, win, hop, 40, 13, 0.0, 8000.0, flags=fl))
        for want in (dict(), dict(want_mfcc=True), dict(want_mfcc=True, want_delta=True)):
            for _ in range(3):
                plan.mfcc_change(pcm, prm, **want)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.iters):
                plan.mfcc_change(pcm, prm, **want)
            e1.record()
            torch.cuda.synchronize()
            print(f"flags {fl} {want or 'totChange only'}: {e0.elapsed_time(e1) / a.iters:.4f} ms")


if __name__ == "__main__":
    main()
