#!/usr/bin/env python
"""Whole-path timing of the other BASELINE configs (parity-test cases, not bench lines).

    python tools/bench_cfgs.py [cfg3] [cfg4] [gui]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import modulation_mfcc_b200 as mm

CFG = {
    # name: (clips, sr, seconds, kwargs)
    "cfg2": (1024, 16000, 10.0, dict(tStep=0.01, winLen=0.025, n_fft=512, n_mels=40, n_mfcc=13)),
    "cfg3": (512, 44100, 10.0, dict(tStep=0.01, winLen=0.025, n_fft=2048, n_mels=128, n_mfcc=20)),
    "cfg4": (1, 16000, 3600.0, dict(tStep=0.01, winLen=0.025, n_fft=512, n_mels=40, n_mfcc=13, mod_hop_s=0.5)),
    "cfg4_slide": (1, 16000, 600.0, dict(tStep=0.01, winLen=0.025, n_fft=512, n_mels=40, n_mfcc=13, mod_hop_s=0.01)),
    "gui": (256, 10000, 10.0, dict(tStep=0.005, winLen=0.025, n_fft=512, n_mels=128, n_mfcc=13, fmin=100.0, fmax=10000.0)),
}


def main():
    dev = torch.device("cuda", 0)
    for name in sys.argv[1:] or ["cfg3", "cfg4"]:
        clips, sr, secs, kw = CFG[name]
        n = int(sr * secs)
        pcm = mm.synth_batch_device(clips, n, sr, seed=7, device=dev)
        fx = mm.FeatureExtractor(sr, device=0, **kw)
        lib = mm.lib()
        for _ in range(2):
            res = fx(pcm, want_logmel=False)
        torch.cuda.synchronize()
        lib.mmf_launch_count(1)
        iters = 5
        t0 = time.perf_counter()
        for _ in range(iters):
            res = fx(pcm, want_logmel=False)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / iters * 1e3
        print(f"{name}: {clips} x {secs:.0f} s @ {sr} Hz: {ms:.3f} ms per pass, {clips * secs / (ms * 1e-3):.3e} audio-s/s, "
              f"{lib.mmf_launch_count(0) // iters} launches, shapes " + ", ".join(f"{k}{tuple(v.shape)}" for k, v in res.items() if hasattr(v, "shape")))
        del pcm, res
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
