"""Mint the golden vectors under tests/golden/ from the CPU oracle.

The reference repository ships no tests, golden vectors or fixtures for this path
and cannot be imported in this image, so these
vectors pin the *oracle* (and, through the GPU parity tests, the CUDA path): one
seed-fixed synthetic clip per BASELINE.json config, shortened so each file stays
small, plus the edge cases of SURVEY.md section 8c.

    python tests/golden/make_golden.py          # rewrites tests/golden/*.npz
"""

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import oracle  # noqa: E402
from modulation_mfcc_b200.synth import synth_clip  # noqa: E402

# name: (seed, sr, seconds, kwargs of oracle.mfcc_features)
CASES = {
    "cfg1_16k_40mel": (0, 16000, 2.0, dict(tStep=0.01, winLen=0.025, n_fft=512, n_mels=40, n_mfcc=13)),
    "cfg3_44k_128mel": (1, 44100, 1.0, dict(tStep=0.01, winLen=0.025, n_fft=2048, n_mels=128, n_mfcc=20)),
    "gui_default_10k": (2, 10000, 2.0, dict(tStep=0.005, winLen=0.025, n_fft=512, n_mels=128, n_mfcc=13, fmin=100, fmax=10000)),
    "cfg4_long_hop": (3, 16000, 4.0, dict(tStep=0.01, winLen=0.025, n_fft=512, n_mels=40, n_mfcc=13, mod_hop_s=0.01)),
}


def edge_clips(n=8000, sr=16000):
    t = np.arange(n) / sr
    z = np.zeros(n, np.float32)
    dc = np.full(n, 0.25, np.float32)
    tone = (0.5 * np.sin(2 * np.pi * (sr / 512 * 20) * t)).astype(np.float32)
    imp0 = z.copy()
    imp0[0] = 1.0
    impN = z.copy()
    impN[-1] = 1.0
    return {"zero": z, "dc": dc, "tone_bin20": tone, "impulse_first": imp0, "impulse_last": impN}


def main():
    for name, (seed, sr, secs, kw) in CASES.items():
        y = synth_clip(seed, int(sr * secs), sr)
        f = oracle.mfcc_features(y, sr, **kw)
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            seed=seed,
            sr=sr,
            n=len(y),
            y_head=y[:64],
            logmel=f["logmel"],
            mfcc=f["mfcc"],
            delta=f["delta"],
            totChange=f["totChange"],
            T=f["T"],
            modspec=f["modspec"][:, :: (10 if name == "cfg4_long_hop" else 1)],
            band_energy=f["band_energy"],
            power_cols=f["power"][:, ::25],
        )
    out = {}
    for k, y in edge_clips().items():
        M, inter = oracle.mfcc(y, 16000, n_mfcc=13, win_length=400, hop_length=160, n_fft=512, fmin=0, fmax=8000, n_mels=40, return_intermediates=True)
        out[k + "_mfcc"] = M
        out[k + "_logmel"] = inter["logmel"]
    # shortest legal clip (T = 22 frames) for get_MFCCS_change
    y22 = synth_clip(9, 21 * 50, 10000)
    kw = dict(tStep=0.005, winLen=0.025, n_mfcc=13, n_fft=512, minFreq=100, maxFreq=10000, outFiltCutOff=[12])
    tot, T = oracle.get_MFCCS_change(y22, 10000, **kw)
    out["t22_tot"], out["t22_T"] = tot, T
    np.savez_compressed(os.path.join(HERE, "edge_cases.npz"), **out)
    print("wrote", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
