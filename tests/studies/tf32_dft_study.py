#!/usr/bin/env python
"""Numerical study for the round-2 tensor-core transform.

A 512-point real frame = 256-point complex FFT of the packed samples = two radix-16 stages, each a real
[rows x 32] . [32 x 32] GEMM.  `tcgen05.mma kind::tf32` reads 10-bit mantissas, so fp32 operands must be
split (x = hi + lo [+ lo2]) and the product rebuilt from several MMAs.  This script emulates that on the
host (TF32 operands = fp32 with the low 13 mantissa bits dropped, products exact, fp32 accumulation) and
reports the error of the linear mel power per band against float64, next to a plain fp32 FFT -- the
quantity the parity tests bound by 1e-4 (tests/test_gpu_parity.py LOGMEL_REL).  `kind::f16` runs at twice the
TF32 rate with the same 11-bit significand; its 5-bit exponent needs a power-of-two scale per frame and the
low parts carried at 2^11 in their own accumulator ("fp16 x3").

    python tests/studies/tf32_dft_study.py            # CPU only, a few seconds
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import mfcc_oracle as oracle  # noqa: E402  (host-side study, not a product path)
from modulation_mfcc_b200.synth import synth_clip  # noqa: E402


def tf32(x):
    u = np.asarray(x, dtype=np.float32).view(np.uint32) & np.uint32(0xFFFFE000)
    return u.view(np.float32)


def split(x, terms):
    """x (fp32) -> list of TF32-representable fp32 arrays whose sum approximates x."""
    out, r = [], np.asarray(x, dtype=np.float32)
    for _ in range(terms):
        h = tf32(r)
        out.append(h)
        r = (r - h).astype(np.float32)  # exact in fp32
    return out


def mma(a_terms, b_terms, pairs):
    """sum over (i, j) in pairs of a_i . b_j, K accumulated in fp32 in slabs of 8 as the tensor core does."""
    rows, K = a_terms[0].shape
    acc = np.zeros((rows, b_terms[0].shape[1]), dtype=np.float32)
    for i, j in pairs:
        a, b = a_terms[i].astype(np.float64), b_terms[j].astype(np.float64)
        for k0 in range(0, K, 8):
            acc = (acc + (a[:, k0:k0 + 8] @ b[k0:k0 + 8, :]).astype(np.float32)).astype(np.float32)
    return acc


def split_f16(x, scale):
    """x (fp32) -> fp16 pair (hi, lo) with x ~= (hi + lo * 2^-11) / scale; values as fp32 arrays."""
    xs = (np.asarray(x, dtype=np.float32) * np.float32(scale)).astype(np.float32)
    hi = xs.astype(np.float16).astype(np.float32)
    lo = ((xs - hi) * np.float32(2048.0)).astype(np.float16).astype(np.float32)
    return hi, lo


def mma_f16x3(a, b_hi, b_lo, row_scale):
    """fp16 x3: acc0 = hi.F_hi, acc1 = hi.F_lo + lo.F_hi (both carry 2^-11), K = 16 slabs, fp32 accumulate."""
    hi, lo = split_f16(a, 1.0)  # caller pre-scales rows
    acc0 = np.zeros((a.shape[0], b_hi.shape[1]), dtype=np.float32)
    acc1 = np.zeros_like(acc0)
    for k0 in range(0, a.shape[1], 16):
        sl = slice(k0, k0 + 16)
        acc0 = (acc0 + (hi[:, sl].astype(np.float64) @ b_hi[sl].astype(np.float64)).astype(np.float32)).astype(np.float32)
        acc1 = (acc1 + (hi[:, sl].astype(np.float64) @ b_lo[sl].astype(np.float64)).astype(np.float32)).astype(np.float32)
        acc1 = (acc1 + (lo[:, sl].astype(np.float64) @ b_hi[sl].astype(np.float64)).astype(np.float32)).astype(np.float32)
    return ((acc0 + acc1 * np.float32(2.0 ** -11)) / row_scale).astype(np.float32)


def fft256_two_stage_f16(z):
    """Same two GEMM stages with fp16 operands: every 8-frame row block is scaled by a power of two so that
    its largest |value| sits near 2^6 (stage outputs grow 16x per stage and must stay below 65504)."""
    F16 = np.exp(-2j * np.pi * np.outer(np.arange(16), np.arange(16)) / 16)
    R = real_rep(F16).astype(np.float32)
    b_hi, b_lo = split_f16(R, 1.0)
    nf = z.shape[0]

    def stage(a_ri):
        rows = a_ri.reshape(nf, -1)
        mx = np.maximum(np.abs(rows).max(axis=1), 1e-30)
        sc = np.exp2(6 - np.ceil(np.log2(mx))).astype(np.float32)  # per frame, power of two
        sc_rows = np.repeat(sc, 16)[:, None]
        return mma_f16x3((a_ri * sc_rows).astype(np.float32), b_hi, b_lo, sc_rows)

    x = z.reshape(nf, 16, 16)
    a = np.transpose(x, (0, 2, 1)).reshape(nf * 16, 16)
    a_ri = np.empty((nf * 16, 32), dtype=np.float32)
    a_ri[:, 0::2], a_ri[:, 1::2] = a.real, a.imag
    d1 = stage(a_ri)
    y = (d1[:, 0::2] + 1j * d1[:, 1::2]).astype(np.complex64).reshape(nf, 16, 16)
    tw = np.exp(-2j * np.pi * np.outer(np.arange(16), np.arange(16)) / 256).astype(np.complex64)
    y = (y * tw[None]).astype(np.complex64)
    a2 = np.transpose(y, (0, 2, 1)).reshape(nf * 16, 16)
    a2_ri = np.empty((nf * 16, 32), dtype=np.float32)
    a2_ri[:, 0::2], a2_ri[:, 1::2] = a2.real, a2.imag
    d2 = stage(a2_ri)
    Z = (d2[:, 0::2] + 1j * d2[:, 1::2]).reshape(nf, 16, 16)
    return np.transpose(Z, (0, 2, 1)).reshape(nf, 256)


def real_rep(F):
    """complex [n x n] -> real [2n x 2n] acting on interleaved (re, im) row vectors from the right."""
    n = F.shape[0]
    R = np.zeros((2 * n, 2 * n))
    R[0::2, 0::2] = F.real
    R[0::2, 1::2] = F.imag
    R[1::2, 0::2] = -F.imag
    R[1::2, 1::2] = F.real
    return R


PAIRS = {
    "tf32 x1": (1, [(0, 0)]),
    "tf32 x3": (2, [(0, 0), (1, 0), (0, 1)]),
    "tf32 x4": (2, [(0, 0), (1, 0), (0, 1), (1, 1)]),
    "tf32 x6": (3, [(0, 0), (1, 0), (0, 1), (1, 1), (2, 0), (0, 2)]),
}


def fft256_two_stage(z, scheme):
    """z: [frames, 256] complex64 -> 256-point FFT through two emulated radix-16 GEMM stages."""
    terms, pairs = PAIRS[scheme]
    F16 = np.exp(-2j * np.pi * np.outer(np.arange(16), np.arange(16)) / 16)
    B = split(real_rep(F16).astype(np.float32), terms)
    nf = z.shape[0]
    # n = n1 + 16 n2, k = 16 k1 + k2:  stage 1 over n2 (-> k2), twiddle W256^(n1 k2), stage 2 over n1 (-> k1)
    x = z.reshape(nf, 16, 16)  # [f, n2, n1]
    a = np.transpose(x, (0, 2, 1)).reshape(nf * 16, 16)  # rows (f, n1), cols n2
    a_ri = np.empty((nf * 16, 32), dtype=np.float32)
    a_ri[:, 0::2], a_ri[:, 1::2] = a.real, a.imag
    d1 = mma(split(a_ri, terms), B, pairs)  # rows (f, n1), cols (k2, re/im)
    y = (d1[:, 0::2] + 1j * d1[:, 1::2]).astype(np.complex64).reshape(nf, 16, 16)  # [f, n1, k2]
    tw = np.exp(-2j * np.pi * np.outer(np.arange(16), np.arange(16)) / 256).astype(np.complex64)
    y = (y * tw[None]).astype(np.complex64)
    a2 = np.transpose(y, (0, 2, 1)).reshape(nf * 16, 16)  # rows (f, k2), cols n1
    a2_ri = np.empty((nf * 16, 32), dtype=np.float32)
    a2_ri[:, 0::2], a2_ri[:, 1::2] = a2.real, a2.imag
    d2 = mma(split(a2_ri, terms), B, pairs)  # rows (f, k2), cols (k1, re/im)
    Z = (d2[:, 0::2] + 1j * d2[:, 1::2]).reshape(nf, 16, 16)  # [f, k2, k1]
    return np.transpose(Z, (0, 2, 1)).reshape(nf, 256)  # k = 16 k1 + k2


def power_from_packed(Z):
    """real-FFT split step in float64 (its fp32 cost is the same for every scheme)."""
    Z = Z.astype(np.complex128)
    k = np.arange(257)
    Zk = Z[:, k % 256]
    Zc = np.conj(Z[:, (256 - k) % 256])
    X = 0.5 * (Zk + Zc) - 0.5j * np.exp(-2j * np.pi * k / 512) * (Zk - Zc)
    return X.real ** 2 + X.imag ** 2


def study(n_clips=4, seconds=2.0, verbose=True):
    """-> {scheme: max relative error of the linear mel power}"""
    sr, n_fft, win, hop, n_mels = 16000, 512, 400, 160, 40
    y = np.stack([synth_clip(s, int(sr * seconds), sr) for s in range(n_clips)])
    w = oracle.padded_hann(win, n_fft).astype(np.float32)
    frames = []
    for clip in y:
        p = np.pad(clip.astype(np.float32), n_fft // 2)
        idx = np.arange(0, len(p) - n_fft + 1, hop)[:, None] + np.arange(n_fft)[None]
        frames.append((p[idx] * w[None]).astype(np.float32))
    fr = np.concatenate(frames)  # [frames, 512] fp32: the same input for every scheme
    mel = oracle.mel_filterbank(sr, n_fft, n_mels, 0.0, sr / 2)
    ref_pow = np.abs(np.fft.rfft(fr.astype(np.float64), axis=1)) ** 2
    ref_mel = ref_pow @ mel.T.astype(np.float64)
    z = (fr[:, 0::2] + 1j * fr[:, 1::2]).astype(np.complex64)
    out = {}

    def report(name, powr):
        m = powr @ mel.T.astype(np.float64)
        rel = np.abs(m - ref_mel) / np.maximum(ref_mel, 1e-300)
        peak = ref_pow.max(axis=1, keepdims=True)
        out[name] = float(rel.max())
        if verbose:
            print(f"{name:>10}: mel power rel err max {rel.max():.2e}  p99.9 {np.quantile(rel, 0.999):.2e}  "
                  f"median {np.median(rel):.2e};  |P - ref| / frame peak max {np.max(np.abs(powr - ref_pow) / peak):.2e}")

    import scipy.fft

    report("fp32 FFT", np.abs(scipy.fft.rfft(fr, axis=1).astype(np.complex128)) ** 2)
    for scheme in PAIRS:
        report(scheme, power_from_packed(fft256_two_stage(z, scheme)))
    report("fp16 x3", power_from_packed(fft256_two_stage_f16(z)))
    return out


def main():
    study()


if __name__ == "__main__":
    main()
