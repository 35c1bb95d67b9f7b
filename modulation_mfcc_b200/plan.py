"""Plans: a frozen frame/spectrum configuration bound to one CUDA device.

A :class:`Plan` owns the device-side constant tables (Hann window, FFT twiddles,
sparse Slaney mel bank, padded DCT rows) and exposes one method per C-ABI entry
point, taking and returning ``torch`` CUDA tensors (torch is used for device
memory and streams only; all arithmetic happens in ``libmmf_b200.so``).
"""

from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from dataclasses import dataclass, replace

import numpy as np

from . import _lib
from ._lib import MmfError, check, mmf_change_params, mmf_config, mmf_modspec_params


@dataclass(frozen=True)
class MfccConfig:
    """Arguments librosa.feature.mfcc sees at script/mfcc.py:387 (+ librosa defaults)."""

    sample_rate: float
    n_fft: int = 512
    win_length: int = 400
    hop_length: int = 160
    n_mels: int = 128
    n_mfcc: int = 13
    fmin: float = 0.0
    fmax: float | None = None
    amin: float = 1e-10
    top_db: float | None = 80.0
    preemph: float = 0.0
    device: int = 0
    flags: int = 0

    def to_c(self) -> mmf_config:
        return mmf_config(
            float(self.sample_rate),
            int(self.n_fft),
            int(self.win_length),
            int(self.hop_length),
            int(self.n_mels),
            int(self.n_mfcc),
            float(self.fmin),
            float(self.sample_rate / 2 if self.fmax is None else self.fmax),
            float(self.amin),
            float(-1.0 if self.top_db is None else self.top_db),
            float(self.preemph),
            int(self.device),
            int(self.flags),
        )

    @property
    def n_bins(self) -> int:
        return self.n_fft // 2 + 1


def frame_sizes(sigSr: float, winLen: float, tStep: float) -> tuple[int, int]:
    """``int(winLen*sigSr)``, ``int(tStep*sigSr)`` -- Python truncation, script/mfcc.py:382-384."""
    return int(winLen * sigSr), int(tStep * sigSr)


def num_frames(n_samples: int, n_fft: int, hop_length: int) -> int:
    return int(_lib.lib().mmf_num_frames(int(n_samples), int(n_fft), int(hop_length)))


def host_tables(cfg: MfccConfig):
    """(window[n_fft], mel[n_mels, F], dct[n_mfcc, n_mels]) exactly as a plan uploads them (no GPU needed)."""
    c = cfg.to_c()
    w = np.zeros(cfg.n_fft, np.float32)
    m = np.zeros((cfg.n_mels, cfg.n_bins), np.float32)
    d = np.zeros((cfg.n_mfcc, cfg.n_mels), np.float32)
    check(_lib.lib().mmf_host_tables(C.byref(c), w.ctypes.data, m.ctypes.data, d.ctypes.data))
    return w, m, d


def sos_zi(sos: np.ndarray):
    """(zi[n_sections, 2], padlen) as scipy.signal.sosfilt_zi / sosfiltfilt compute them."""
    sos = np.ascontiguousarray(sos, dtype=np.float64)
    zi = np.zeros((sos.shape[0], 2), np.float64)
    padlen = C.c_int32(0)
    check(_lib.lib().mmf_sos_zi(sos.ctypes.data, sos.shape[0], zi.ctypes.data, C.byref(padlen)))
    return zi, int(padlen.value)


def make_change_params(
    sos: np.ndarray, *, remove_first: int = 1, diff_method: int = 0, out_sos: np.ndarray | None = None
) -> mmf_change_params:
    """Pack the post-MFCC parameters of get_MFCCS_change (script/mfcc.py:393-425)."""
    prm = mmf_change_params()
    sos = np.ascontiguousarray(sos, dtype=np.float64)
    if sos.ndim != 2 or sos.shape[1] != 6 or sos.shape[0] > 16:
        raise ValueError("sos must be [n_sections <= 16, 6]")
    prm.remove_first = int(bool(remove_first))
    prm.diff_method = int(diff_method)
    prm.n_sections = sos.shape[0]
    for i, v in enumerate(sos.ravel()):
        prm.sos[i] = v
    if out_sos is None:
        prm.out_kind = 1
        prm.out_n_sections = 0
    else:
        out_sos = np.ascontiguousarray(out_sos, dtype=np.float64)
        if out_sos.ndim != 2 or out_sos.shape[1] != 6 or out_sos.shape[0] > 16:
            raise ValueError("out_sos must be [n_sections <= 16, 6]")
        prm.out_kind = 0
        prm.out_n_sections = out_sos.shape[0]
        for i, v in enumerate(out_sos.ravel()):
            prm.out_sos[i] = v
    return prm


def _torch():
    import torch

    return torch


def _stream_ptr(device) -> int:
    return int(_torch().cuda.current_stream(device).cuda_stream)


def design_resample_filter(n_in: int, up: int, down: int, quality: str = "scipy"):
    """Taps (float32, zero-padded for the polyphase bookkeeping), samples to drop at the front of the full
    convolution, and output length, for ``y = upfirdn(h, x, up, down)[n_pre_remove : n_pre_remove + n_out]``.

    ``"scipy"`` restates scipy/signal/_signaltools.py resample_poly (filter design and edge bookkeeping) -- the
    output is bit-compatible with it up to float32 rounding.  ``"hq"`` keeps the bookkeeping and swaps the filter:
    Kaiser-windowed sinc with the transition band 0.913 .. 1.0 of the lower Nyquist at 120 dB (beta 12.27,
    90 taps per phase side), i.e. libsoxr HQ's band limits -- what ``librosa.load`` applies at script/mfcc.py:373.
    Content below 0.9 of the lower Nyquist comes through within 1e-5 of the ideal band-limited resampling
    (tests/test_host.py::test_hq_resampler_is_transparent_below_the_band_edge); libsoxr's own output cannot be
    compared here (not installed, no network)."""
    import scipy.signal
    from scipy.signal._upfirdn import _output_len

    n_out = n_in * up
    n_out = n_out // down + bool(n_out % down)
    max_rate = max(up, down)
    if quality == "scipy":
        half_len = 10 * max_rate
        h = scipy.signal.firwin(2 * half_len + 1, 1.0 / max_rate, window=("kaiser", 5.0))
    elif quality == "hq":
        half_len = 90 * max_rate
        h = scipy.signal.firwin(2 * half_len + 1, 0.9565 / max_rate, window=("kaiser", 12.27))
    else:
        raise ValueError("quality must be 'scipy' or 'hq'")
    h = h.astype(np.float32) * up
    n_pre_pad = down - half_len % down
    n_post_pad = 0
    n_pre_remove = (half_len + n_pre_pad) // down
    while _output_len(len(h) + n_pre_pad + n_post_pad, n_in, up, down) < n_out + n_pre_remove:
        n_post_pad += 1
    h = np.concatenate((np.zeros(n_pre_pad, dtype=h.dtype), h, np.zeros(n_post_pad, dtype=h.dtype)))
    return np.ascontiguousarray(h, dtype=np.float32), n_pre_remove, n_out


class Plan:
    """Device-resident plan (``mmf_plan``).  Not thread-safe: one plan per host thread."""

    def __init__(self, cfg: MfccConfig):
        torch = _torch()
        if not torch.cuda.is_available():
            raise MmfError(_lib.MMF_ERR_CUDA, "no CUDA device available (this package has no CPU fallback)")
        self.cfg = cfg
        self.device = torch.device("cuda", cfg.device)
        self._h = C.c_void_p()
        c = cfg.to_c()
        check(_lib.lib().mmf_plan_create(C.byref(self._h), C.byref(c)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            _lib.lib().mmf_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers -----------------------------------------------------------
    def _pcm(self, pcm):
        torch = _torch()
        if not isinstance(pcm, torch.Tensor):
            pcm = torch.as_tensor(np.ascontiguousarray(pcm, dtype=np.float32)).to(self.device, non_blocking=True)
        if pcm.device != self.device:
            pcm = pcm.to(self.device)
        if pcm.dtype != torch.float32:
            pcm = pcm.to(torch.float32)
        if pcm.dim() == 1:
            pcm = pcm[None, :]
        if pcm.dim() != 2:
            raise ValueError("pcm must be [n_clips, n_samples]")
        if pcm.stride(1) != 1:
            pcm = pcm.contiguous()
        return pcm

    def num_frames(self, n_samples: int) -> int:
        return num_frames(n_samples, self.cfg.n_fft, self.cfg.hop_length)

    # -- kernels -----------------------------------------------------------
    def stft_power(self, pcm):
        """|STFT|^2 -> float32 [B, F, T]."""
        torch = _torch()
        pcm = self._pcm(pcm)
        B, N = pcm.shape
        T = self.num_frames(N)
        out = torch.empty((B, self.cfg.n_bins, T), device=self.device, dtype=torch.float32)
        check(
            _lib.lib().mmf_stft_power(
                self._h, pcm.data_ptr(), B, N, pcm.stride(0) if B > 1 else N, out.data_ptr(), _stream_ptr(self.device)
            )
        )
        return out

    def logmel(self, pcm):
        """Unclamped log-mel [B, n_mels, T] float32 and the per-clip max keys [B] int32."""
        torch = _torch()
        pcm = self._pcm(pcm)
        B, N = pcm.shape
        T = self.num_frames(N)
        out = torch.empty((B, self.cfg.n_mels, T), device=self.device, dtype=torch.float32)
        cmax = torch.empty((B,), device=self.device, dtype=torch.int32)
        check(
            _lib.lib().mmf_logmel(
                self._h,
                pcm.data_ptr(),
                B,
                N,
                pcm.stride(0) if B > 1 else N,
                out.data_ptr(),
                cmax.data_ptr(),
                _stream_ptr(self.device),
            )
        )
        return out, cmax

    def mfcc(self, logmel, clipmax, *, delta: bool = False, clamp_in_place: bool = True):
        """top_db clamp + DCT-II: MFCC [B, n_mfcc, T] (and np.gradient delta)."""
        torch = _torch()
        B, _, T = logmel.shape
        out = torch.empty((B, self.cfg.n_mfcc, T), device=self.device, dtype=torch.float32)
        d = torch.empty_like(out) if delta else None
        check(
            _lib.lib().mmf_mfcc(
                self._h,
                logmel.data_ptr(),
                clipmax.data_ptr(),
                B,
                T,
                out.data_ptr(),
                d.data_ptr() if delta else None,
                1 if clamp_in_place else 0,
                _stream_ptr(self.device),
            )
        )
        return (out, d) if delta else out

    def sosfiltfilt(self, x, sos):
        """scipy.signal.sosfiltfilt(sos, x) along the last axis -> float64."""
        torch = _torch()
        sos = np.ascontiguousarray(sos, dtype=np.float64)
        if x.dtype not in (torch.float32, torch.float64):
            x = x.to(torch.float64)
        x = x.contiguous()
        T = x.shape[-1]
        rows = x.numel() // T
        y = torch.empty(x.shape, device=self.device, dtype=torch.float64)
        check(
            _lib.lib().mmf_sosfiltfilt(
                self._h,
                x.data_ptr(),
                1 if x.dtype == torch.float32 else 0,
                rows,
                T,
                T,
                sos.ctypes.data,
                sos.shape[0],
                y.data_ptr(),
                T,
                _stream_ptr(self.device),
            )
        )
        return y

    def delta_norm(self, x, method: int = 0):
        """x float64 [B, rows, T] -> sqrt(sum_rows d^2)/rows, float64 [B, T]."""
        torch = _torch()
        x = x.contiguous()
        B, R, T = x.shape
        out = torch.empty((B, T), device=self.device, dtype=torch.float64)
        check(_lib.lib().mmf_delta_norm(self._h, x.data_ptr(), B, R, T, method, out.data_ptr(), _stream_ptr(self.device)))
        return out

    def fir_filtfilt(self, x, b):
        """scipy.signal.filtfilt(b, 1, x) along the last axis, float64."""
        torch = _torch()
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = x.to(torch.float64).contiguous()
        T = x.shape[-1]
        rows = x.numel() // T
        y = torch.empty_like(x)
        work = torch.empty((rows, T + 6 * len(b)), device=self.device, dtype=torch.float64)
        check(
            _lib.lib().mmf_fir_filtfilt(
                self._h, x.data_ptr(), rows, T, b.ctypes.data, len(b), y.data_ptr(), work.data_ptr(), _stream_ptr(self.device)
            )
        )
        return y

    def stencil(self, x, coef, edge_l, edge_r):
        """Banded stencil with dense boundary rows along the last axis, float64."""
        torch = _torch()
        coef = np.ascontiguousarray(coef, dtype=np.float64)
        edge_l = np.ascontiguousarray(edge_l, dtype=np.float64)
        edge_r = np.ascontiguousarray(edge_r, dtype=np.float64)
        half = (len(coef) - 1) // 2
        n_edge, n_edge_in = edge_l.shape if edge_l.size else (0, 0)
        x = x.to(torch.float64).contiguous()
        T = x.shape[-1]
        rows = x.numel() // T
        y = torch.empty_like(x)
        check(
            _lib.lib().mmf_stencil(
                self._h,
                x.data_ptr(),
                rows,
                T,
                coef.ctypes.data,
                half,
                edge_l.ctypes.data if n_edge else None,
                edge_r.ctypes.data if n_edge else None,
                n_edge,
                n_edge_in,
                y.data_ptr(),
                _stream_ptr(self.device),
            )
        )
        return y

    def modspec(self, mfcc, win: int, hop: int, nfft: int, band_bins=None, *, want_mag: bool = True):
        """Modulation spectrum of [B, C, T] float32 trajectories.

        Returns (mag [B, C, n_win, nfft/2+1] or None, band [B, n_win, n_bands] or None)."""
        torch = _torch()
        mfcc = mfcc.contiguous()
        B, Cc, T = mfcc.shape
        n_win = 1 + (T - win) // hop if T >= win else 0
        nb = nfft // 2 + 1
        mag = torch.empty((B, Cc, n_win, nb), device=self.device, dtype=torch.float32) if want_mag else None
        band = None
        lo = hi = None
        n_bands = 0
        if band_bins is not None and len(band_bins):
            n_bands = len(band_bins)
            lo = np.ascontiguousarray([b[0] for b in band_bins], dtype=np.int32)
            hi = np.ascontiguousarray([b[1] for b in band_bins], dtype=np.int32)
            band = torch.empty((B, n_win, n_bands), device=self.device, dtype=torch.float32)
        if n_win > 0:
            check(
                _lib.lib().mmf_modspec(
                    self._h,
                    mfcc.data_ptr(),
                    B,
                    Cc,
                    T,
                    win,
                    hop,
                    nfft,
                    mag.data_ptr() if want_mag else None,
                    band.data_ptr() if band is not None else None,
                    lo.ctypes.data if lo is not None else None,
                    hi.ctypes.data if hi is not None else None,
                    n_bands,
                    _stream_ptr(self.device),
                )
            )
        return mag, band

    def rms(self, pcm, frame_length: int, hop_length: int, center: bool = True):
        """librosa.feature.rms -> float32 [B, T_rms]."""
        torch = _torch()
        pcm = self._pcm(pcm)
        B, N = pcm.shape
        pad = frame_length // 2 if center else 0
        if N + 2 * pad < frame_length:
            raise ValueError(f"Input is too short (n={N}) for frame_length={frame_length}")
        T = 1 + (N + 2 * pad - frame_length) // hop_length
        out = torch.empty((B, T), device=self.device, dtype=torch.float32)
        check(
            _lib.lib().mmf_rms(
                self._h,
                pcm.data_ptr(),
                B,
                N,
                pcm.stride(0) if B > 1 else N,
                frame_length,
                hop_length,
                1 if center else 0,
                out.data_ptr(),
                _stream_ptr(self.device),
            )
        )
        return out

    def hilbert_envelope(self, x):
        """``np.abs(scipy.signal.hilbert(x))`` along the last axis, float32 on the device."""
        torch = _torch()
        x = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32))
        x = x.to(device=torch.device("cuda", self.cfg.device), dtype=torch.float32)
        squeeze = x.ndim == 1
        if squeeze:
            x = x[None, :]
        x = x.contiguous()
        amp = torch.empty_like(x)
        check(
            _lib.lib().mmf_hilbert_envelope(
                self._h, x.data_ptr(), x.shape[0], x.shape[1], x.stride(0), amp.data_ptr(), amp.stride(0), _stream_ptr(x.device)
            )
        )
        return amp[0] if squeeze else amp

    def find_peaks(self, x, *, minima: bool = False, max_peaks: int | None = None):
        """``scipy.signal.find_peaks(x)`` (default arguments) for every float64 row of ``x`` on the
        device.  Returns ``(idx [rows, max_peaks] int32, count [rows] int32)``; row ``r`` holds its
        ``min(count[r], max_peaks)`` peak indices in ascending order."""
        torch = _torch()
        x = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64))
        x = x.to(device=torch.device("cuda", self.cfg.device), dtype=torch.float64)
        if x.ndim == 1:
            x = x[None, :]
        x = x.contiguous()
        rows, T = x.shape
        if max_peaks is None:
            max_peaks = max(1, (T - 1) // 2)  # peaks need a lower sample between them
        idx = torch.full((rows, max_peaks), -1, device=x.device, dtype=torch.int32)
        cnt = torch.empty((rows,), device=x.device, dtype=torch.int32)
        check(
            _lib.lib().mmf_find_peaks(
                self._h, x.data_ptr(), rows, T, x.stride(0), 1 if minima else 0, max_peaks, idx.data_ptr(), cnt.data_ptr(),
                _stream_ptr(x.device),
            )
        )
        return idx, cnt

    def pcm16_to_f32(self, pcm16):
        """int16 PCM (CUDA tensor or numpy) -> float32 in [-1, 1) on the device (x / 32768)."""
        torch = _torch()
        x = pcm16 if isinstance(pcm16, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(pcm16))
        x = x.to(device=torch.device("cuda", self.cfg.device), dtype=torch.int16).contiguous()
        y = torch.empty(x.shape, device=x.device, dtype=torch.float32)
        check(_lib.lib().mmf_pcm16_to_f32(self._h, x.data_ptr(), x.numel(), y.data_ptr(), _stream_ptr(x.device)))
        return y

    def resample_poly(self, x, up: int, down: int, quality: str = "scipy"):
        """Rational-rate conversion of float32 rows on the device (polyphase FIR, float64 accumulation).

        ``quality="scipy"``: exactly ``scipy.signal.resample_poly(x, up, down, axis=-1)`` (Kaiser-5 FIR, 10 taps per
        phase side, constant padding).  ``quality="hq"``: the band limits of libsoxr's HQ recipe that
        ``librosa.load(sr=...)`` uses at script/mfcc.py:373 (pass band to 0.913 of the lower Nyquist, stop band from
        the Nyquist, 120 dB).  The filter is designed on the host (:func:`design_resample_filter`)."""
        torch = _torch()
        import math

        x = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32))
        x = x.to(device=torch.device("cuda", self.cfg.device), dtype=torch.float32)
        squeeze = x.ndim == 1
        if squeeze:
            x = x[None, :]
        x = x.contiguous()
        g = math.gcd(int(up), int(down))
        up, down = int(up) // g, int(down) // g
        n_in = x.shape[-1]
        if up == 1 and down == 1:
            return x[0] if squeeze else x
        h, n_pre_remove, n_out = design_resample_filter(n_in, up, down, quality)
        y = torch.empty((x.shape[0], n_out), device=x.device, dtype=torch.float32)
        for c0 in range(0, x.shape[0], 65535):
            xs = x[c0 : c0 + 65535]
            ys = y[c0 : c0 + 65535]
            check(
                _lib.lib().mmf_resample_poly(
                    self._h, xs.data_ptr(), xs.shape[0], n_in, xs.stride(0), h.ctypes.data, len(h), up, down,
                    n_pre_remove, n_out, ys.data_ptr(), ys.stride(0), _stream_ptr(x.device),
                )
            )
        return y[0] if squeeze else y

    def mfcc_change(self, pcm, prm: mmf_change_params, *, want_logmel=False, want_mfcc=False, want_delta=False):
        """Whole get_MFCCS_change on device-resident PCM -> dict of CUDA tensors."""
        torch = _torch()
        pcm = self._pcm(pcm)
        B, N = pcm.shape
        T = self.num_frames(N)
        tot = torch.empty((B, T), device=self.device, dtype=torch.float64)
        logmel = torch.empty((B, self.cfg.n_mels, T), device=self.device, dtype=torch.float32) if want_logmel else None
        mfcc = (
            torch.empty((B, self.cfg.n_mfcc, T), device=self.device, dtype=torch.float32)
            if (want_mfcc or want_delta)
            else None
        )
        delta = torch.empty((B, self.cfg.n_mfcc, T), device=self.device, dtype=torch.float32) if want_delta else None
        check(
            _lib.lib().mmf_mfcc_change(
                self._h,
                pcm.data_ptr(),
                B,
                N,
                pcm.stride(0) if B > 1 else N,
                C.byref(prm),
                tot.data_ptr(),
                logmel.data_ptr() if logmel is not None else None,
                mfcc.data_ptr() if mfcc is not None else None,
                delta.data_ptr() if delta is not None else None,
                _stream_ptr(self.device),
            )
        )
        return {"totChange": tot, "logmel": logmel, "mfcc": mfcc, "delta": delta}

    def change_from_logmel(self, logmel, clipmax, prm: mmf_change_params, *, want_delta=True, clamp_in_place=True):
        """Second half of :meth:`mfcc_change` for a log-mel already on the device."""
        torch = _torch()
        B, _, T = logmel.shape
        tot = torch.empty((B, T), device=self.device, dtype=torch.float64)
        mfcc = torch.empty((B, self.cfg.n_mfcc, T), device=self.device, dtype=torch.float32)
        delta = torch.empty_like(mfcc) if want_delta else None
        check(
            _lib.lib().mmf_change_from_logmel(
                self._h,
                logmel.data_ptr(),
                clipmax.data_ptr(),
                B,
                T,
                C.byref(prm),
                tot.data_ptr(),
                mfcc.data_ptr(),
                delta.data_ptr() if delta is not None else None,
                1 if clamp_in_place else 0,
                _stream_ptr(self.device),
            )
        )
        return {"totChange": tot, "mfcc": mfcc, "delta": delta}

    def features_host(self, pcm_host: np.ndarray, prm: mmf_change_params, mod=None, *, want=("totChange",), out=None):
        """The whole bundle through ONE C-ABI call with host buffers (H2D/D2H inside).

        ``mod`` = (win, hop, nfft, band_bins) or None; ``want`` subset of
        totChange/mfcc/delta/modspec/band_energy; ``out`` may hold preallocated
        (ideally pinned) numpy arrays to receive the results."""
        pcm_host = np.asarray(pcm_host)
        pcm16 = pcm_host.dtype == np.int16  # raw WAV samples: scaled by 1/32768 on the device
        if not pcm16 and pcm_host.dtype != np.float32:
            pcm_host = pcm_host.astype(np.float32)
        if pcm_host.ndim == 1:
            pcm_host = pcm_host[None, :]
        isz = pcm_host.dtype.itemsize
        if pcm_host.strides[1] != isz:
            pcm_host = np.ascontiguousarray(pcm_host)
        B, N = pcm_host.shape
        T = self.num_frames(N)
        nm = self.cfg.n_mfcc
        out = {} if out is None else out
        mp = None
        n_win = nb = n_bands = 0
        if mod is not None:
            win, hop, nfft, bins = mod
            mp = mmf_modspec_params()
            mp.win, mp.hop, mp.nfft, mp.n_bands = int(win), int(hop), int(nfft), len(bins)
            for i, (lo, hi) in enumerate(bins):
                mp.band_lo[i], mp.band_hi[i] = int(lo), int(hi)
            n_win = 1 + (T - win) // hop if T >= win else 0
            nb, n_bands = nfft // 2 + 1, len(bins)
        shapes = {
            "totChange": ((B, T), np.float64),
            "mfcc": ((B, nm, T), np.float32),
            "delta": ((B, nm, T), np.float32),
            "modspec": ((B, nm, n_win, nb), np.float32),
            "band_energy": ((B, n_win, n_bands), np.float32),
        }
        ptr = {}
        for k, (shp, dt) in shapes.items():
            if k in want or k == "totChange":
                a = out.get(k)
                if a is None or a.shape != shp or a.dtype != dt or not a.flags.c_contiguous:
                    a = np.empty(shp, dt)
                    out[k] = a
                ptr[k] = a.ctypes.data
            else:
                ptr[k] = None
        fn = _lib.lib().mmf_features_host_pcm16 if pcm16 else _lib.lib().mmf_features_host
        check(
            fn(
                self._h,
                pcm_host.ctypes.data,
                B,
                N,
                pcm_host.strides[0] // isz if B > 1 else N,
                C.byref(prm),
                C.byref(mp) if mp is not None else None,
                ptr["totChange"],
                ptr["mfcc"],
                ptr["delta"],
                ptr["modspec"],
                ptr["band_energy"],
            )
        )
        return out

    def mfcc_change_host(self, pcm_host: np.ndarray, prm: mmf_change_params, *, want_mfcc: bool = False):
        """Host buffers in and out through the single C-ABI call (H2D/D2H inside).

        float32 samples go through ``mmf_mfcc_change_host``; int16 samples (raw WAV PCM) through
        ``mmf_features_host_pcm16``, which scales them by 1/32768 on the device."""
        pcm_host = np.asarray(pcm_host)
        pcm16 = pcm_host.dtype == np.int16
        if not pcm16 and pcm_host.dtype != np.float32:
            pcm_host = pcm_host.astype(np.float32)
        if pcm_host.ndim == 1:
            pcm_host = pcm_host[None, :]
        isz = pcm_host.dtype.itemsize
        if pcm_host.strides[1] != isz or pcm_host.strides[0] % isz:
            pcm_host = np.ascontiguousarray(pcm_host)
        B, N = pcm_host.shape
        T = self.num_frames(N)
        tot = np.empty((B, T), np.float64)
        mf = np.empty((B, self.cfg.n_mfcc, T), np.float32) if want_mfcc else None
        stride = pcm_host.strides[0] // isz if B > 1 else N
        if pcm16:
            check(
                _lib.lib().mmf_features_host_pcm16(
                    self._h, pcm_host.ctypes.data, B, N, stride, C.byref(prm), None, tot.ctypes.data,
                    mf.ctypes.data if want_mfcc else None, None, None, None,
                )
            )
        else:
            check(
                _lib.lib().mmf_mfcc_change_host(
                    self._h, pcm_host.ctypes.data, B, N, stride, C.byref(prm), tot.ctypes.data,
                    mf.ctypes.data if want_mfcc else None,
                )
            )
        return (tot, mf) if want_mfcc else tot


_PLANS: "OrderedDict[MfccConfig, Plan]" = OrderedDict()
_MAX_PLANS = 16


def get_plan(cfg: MfccConfig) -> Plan:
    """Keyed least-recently-used cache of plans (window, mel bank, DCT, twiddles stay on the device).

    An evicted plan is only dropped from the cache: objects that still hold it (a ``FeatureExtractor``, a
    caller's local) keep a live handle, and ``Plan.__del__`` frees the device memory with the last reference."""
    p = _PLANS.get(cfg)
    if p is None:
        while len(_PLANS) >= _MAX_PLANS:
            _PLANS.popitem(last=False)
        p = Plan(cfg)
        _PLANS[cfg] = p
    else:
        _PLANS.move_to_end(cfg)
    return p


def clear_plans() -> None:
    for p in _PLANS.values():
        p.close()
    _PLANS.clear()


__all__ = [
    "design_resample_filter",
    "MfccConfig",
    "Plan",
    "get_plan",
    "clear_plans",
    "frame_sizes",
    "num_frames",
    "host_tables",
    "sos_zi",
    "make_change_params",
    "replace",
]
