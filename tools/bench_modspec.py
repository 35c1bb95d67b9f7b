#!/usr/bin/env python
"""Time the modulation-spectrum kernel alone on the bench shape (1024 clips x 13 x 1001, win 100, hop 50,
nfft 128): tcgen05 GEMM (flags 0, default) vs FP32 register FFT (MMF_FLAG_NO_TC_MODSPEC = 512).

    python tools/bench_modspec.py [--flags 0,512] [--iters 20]
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
    ap.add_argument("--flags", default="0,512")
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    torch.manual_seed(0)
    M = (torch.randn(a.clips, 13, 1001, device="cuda").cumsum(-1) * 0.5).contiguous()
    Lw, Hw, nfft, n_win = mm.modspec_sizes(1001, 100.0, 1.0, 0.5)
    bins = mm.band_bins(nfft, 100.0)
    for fl in [int(x) for x in a.flags.split(",")]:
        plan = mm.get_plan(mm.MfccConfig(16000, 512, 400, 160, 40, 13, 0.0, 8000.0, flags=fl))
        for _ in range(3):
            plan.modspec(M, Lw, Hw, nfft, bins)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            plan.modspec(M, Lw, Hw, nfft, bins)
        e1.record()
        torch.cuda.synchronize()
        print(f"flags {fl}: {e0.elapsed_time(e1) / a.iters * 1e3:.1f} us per {a.clips} clips")


if __name__ == "__main__":
    main()
