"""CPU restatement of the reference's MFCC-change path (TEST INFRASTRUCTURE ONLY).

See ``oracle/__init__.py`` for the scope statement and the "parity unpinned"
note.  Reference line numbers are relative to ``/root/reference``.

Everything here is plain numpy/scipy, float32 where the reference's librosa
chain is float32 and float64 where scipy promotes (``sosfiltfilt`` on).
"""

from __future__ import annotations

import math

import numpy as np
import scipy.fftpack
import scipy.signal

# ----------------------------------------------------------------------------
# A.1  integer frame sizes  (script/mfcc.py:382-384)
# ----------------------------------------------------------------------------


def frame_sizes(sigSr: float, winLen: float, tStep: float) -> tuple[int, int]:
    """``win_length=int(winLen*sigSr)``, ``hop_length=int(tStep*sigSr)``.

    Python ``int`` truncation of the float product, exactly as
    script/mfcc.py:382 and :384 do (44.1 kHz * 0.025 -> 1102, not 1103).
    """
    return int(winLen * sigSr), int(tStep * sigSr)


# ----------------------------------------------------------------------------
# A.2 / A.3 / A.4  window, framing, spectrum  (librosa.stft as called through
# librosa.feature.mfcc at script/mfcc.py:387)
# ----------------------------------------------------------------------------


def padded_hann(win_length: int, n_fft: int) -> np.ndarray:
    """Periodic Hann of ``win_length`` zero-padded (centred) to ``n_fft``.

    librosa ``stft``: ``get_window('hann', win_length, fftbins=True)`` followed by
    ``util.pad_center(size=n_fft)``; ``lpad = (n_fft - win_length)//2``.  float64,
    as scipy returns it (librosa multiplies the float32 frames by this float64
    window, so the product and the FFT are carried out in float64).
    """
    if win_length > n_fft:
        # librosa.util.pad_center raises ParameterError here
        raise ValueError(
            f"Target size ({n_fft}) must be at least input size ({win_length})"
        )
    n = np.arange(win_length, dtype=np.float64)
    w = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / win_length)
    lpad = (n_fft - win_length) // 2
    out = np.zeros(n_fft, dtype=np.float64)
    out[lpad : lpad + win_length] = w
    return out


def n_frames(n_samples: int, n_fft: int, hop: int) -> int:
    """``1 + (len(y) + 2*(n_fft//2) - n_fft)//hop`` (centre-padded framing)."""
    return 1 + (n_samples + 2 * (n_fft // 2) - n_fft) // hop


def stft_power(
    y: np.ndarray, n_fft: int, hop_length: int, win_length: int, *, chunk: int = 4096
) -> np.ndarray:
    """``np.abs(librosa.stft(y, center=True, pad_mode='constant'))**2`` -> [F, T] f32.

    Frame t is ``y_pad[t*hop : t*hop + n_fft]`` with ``y_pad`` = y padded with
    ``n_fft//2`` zeros on both sides; no pre-emphasis, dither or DC removal.
    The stored spectrum is complex64 (librosa ``dtype_r2c(float32)``), so
    ``abs()**2`` is float32.
    """
    y = np.asarray(y)
    if y.ndim != 1:
        raise ValueError("oracle stft_power expects a 1-D signal")
    w = padded_hann(win_length, n_fft)
    pad = n_fft // 2
    y_pad = np.concatenate([np.zeros(pad, y.dtype), y, np.zeros(pad, y.dtype)])
    T = n_frames(len(y), n_fft, hop_length)
    if T < 1:
        raise ValueError("input too short for one frame")
    frames = np.lib.stride_tricks.as_strided(
        y_pad,
        shape=(T, n_fft),
        strides=(y_pad.strides[0] * hop_length, y_pad.strides[0]),
        writeable=False,
    )
    cdtype = np.complex64 if y.dtype == np.float32 else np.complex128
    F = 1 + n_fft // 2
    out = np.empty((F, T), dtype=np.float32 if cdtype == np.complex64 else np.float64)
    for s in range(0, T, chunk):  # column blocks, like librosa's memory-capped loop
        X = np.fft.rfft(frames[s : s + chunk] * w[None, :], axis=1).astype(cdtype)
        out[:, s : s + chunk] = (np.abs(X) ** 2).T
    return out


# ----------------------------------------------------------------------------
# A.5  Slaney mel filterbank  (librosa.filters.mel, htk=False, norm='slaney')
# ----------------------------------------------------------------------------

_F_SP = 200.0 / 3
_MIN_LOG_HZ = 1000.0
_MIN_LOG_MEL = _MIN_LOG_HZ / _F_SP  # 15.0
_LOGSTEP = math.log(6.4) / 27.0


def hz_to_mel(f):
    f = np.asanyarray(f, dtype=np.float64)
    mels = f / _F_SP
    big = f >= _MIN_LOG_HZ
    with np.errstate(divide="ignore", invalid="ignore"):
        mels = np.where(big, _MIN_LOG_MEL + np.log(np.where(big, f, 1.0) / _MIN_LOG_HZ) / _LOGSTEP, mels)
    return mels


def mel_to_hz(m):
    m = np.asanyarray(m, dtype=np.float64)
    f = _F_SP * m
    big = m >= _MIN_LOG_MEL
    return np.where(big, _MIN_LOG_HZ * np.exp(_LOGSTEP * (m - _MIN_LOG_MEL)), f)


def mel_filterbank(sr: float, n_fft: int, n_mels: int = 128, fmin: float = 0.0, fmax: float | None = None) -> np.ndarray:
    """``librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax, htk=False, norm='slaney')``.

    Returns float32 ``[n_mels, 1 + n_fft//2]``.  ``fmax`` above Nyquist is legal
    (script/main.py:739 passes 10 kHz at sr 10 kHz) and leaves the top filters
    all-zero.  Rounding order follows librosa: triangles are stored into a
    float32 array, then scaled in place by the float64 Slaney norm.
    """
    if fmax is None:
        fmax = float(sr) / 2
    F = 1 + n_fft // 2
    fftfreqs = np.fft.rfftfreq(n=n_fft, d=1.0 / sr)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    weights = np.zeros((n_mels, F), dtype=np.float32)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2 : n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights


# ----------------------------------------------------------------------------
# A.6  power_to_db   (librosa.power_to_db(ref=1.0, amin=1e-10, top_db=80))
# ----------------------------------------------------------------------------


def power_to_db(S: np.ndarray, amin: float = 1e-10, top_db: float | None = 80.0) -> np.ndarray:
    """``10*log10(max(amin, S))`` then clamp to ``max_over_whole_array - top_db``."""
    S = np.asarray(S)
    log_spec = 10.0 * np.log10(np.maximum(amin, S))
    log_spec -= 10.0 * np.log10(np.maximum(amin, 1.0))
    if top_db is not None:
        log_spec = np.maximum(log_spec, log_spec.max() - top_db)
    return log_spec


# ----------------------------------------------------------------------------
# A.7  DCT-II, orthonormal   (scipy.fftpack.dct(type=2, norm='ortho'))
# ----------------------------------------------------------------------------


def dct_ortho_matrix(n_mfcc: int, n_mels: int) -> np.ndarray:
    """float64 ``[n_mfcc, n_mels]`` matrix equivalent to scipy's ortho DCT-II."""
    m = np.arange(n_mels, dtype=np.float64)
    k = np.arange(n_mfcc, dtype=np.float64)[:, None]
    D = np.sqrt(2.0 / n_mels) * np.cos(np.pi * k * (2.0 * m + 1.0) / (2.0 * n_mels))
    D[0, :] = 1.0 / np.sqrt(n_mels)
    return D


def mfcc(
    y: np.ndarray,
    sr: float,
    *,
    n_mfcc: int = 20,
    win_length: int,
    hop_length: int,
    n_fft: int = 2048,
    fmin: float = 0.0,
    fmax: float | None = None,
    n_mels: int = 128,
    top_db: float | None = 80.0,
    amin: float = 1e-10,
    return_intermediates: bool = False,
):
    """``librosa.feature.mfcc(y=y, sr=sr, n_mfcc=, win_length=, hop_length=, n_fft=, fmin=, fmax=)``.

    dct_type=2, norm='ortho', lifter=0; ``n_mels`` defaults to librosa's 128
    because script/mfcc.py:387 never passes it.
    """
    S = stft_power(np.asarray(y), n_fft, hop_length, win_length)
    mel_basis = mel_filterbank(sr, n_fft, n_mels, fmin, fmax)
    melspec = np.einsum("ft,mf->mt", S, mel_basis, optimize=True)
    S_db = power_to_db(melspec, amin=amin, top_db=top_db)
    M = scipy.fftpack.dct(S_db, axis=-2, type=2, norm="ortho")[:n_mfcc, :]
    if return_intermediates:
        return M, {"power": S, "mel_basis": mel_basis, "melspec": melspec, "logmel": S_db}
    return M


# ----------------------------------------------------------------------------
# A.10  zero-phase SOS filter, restated (documentation of what the CUDA scan
# kernel implements; scipy.signal.sosfiltfilt is the authority)
# ----------------------------------------------------------------------------


def sosfiltfilt_restated(sos: np.ndarray, x: np.ndarray) -> np.ndarray:
    """Pure-numpy ``scipy.signal.sosfiltfilt(sos, x)`` along the last axis.

    padlen = 3*(2*n_sections + 1 - min(#(b2==0), #(a2==0))); odd extension;
    ``zi = sosfilt_zi(sos)`` scaled by the first sample of each pass; forward,
    reverse, forward, reverse, trim.  Biquads are direct-form-II transposed.
    """
    sos = np.asarray(sos, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    n_sections = sos.shape[0]
    ntaps = 2 * n_sections + 1
    ntaps -= min((sos[:, 2] == 0).sum(), (sos[:, 5] == 0).sum())
    padlen = 3 * ntaps
    if x.shape[-1] <= padlen:
        raise ValueError(
            "The length of the input vector x must be greater than padlen, which is %d." % padlen
        )
    # odd extension
    left = 2 * x[..., :1] - x[..., padlen:0:-1]
    right = 2 * x[..., -1:] - x[..., -2 : -padlen - 2 : -1]
    ext = np.concatenate([left, x, right], axis=-1)
    # steady-state initial conditions of the cascade for a unit step
    zi = np.zeros((n_sections, 2))
    scale = 1.0
    for s in range(n_sections):
        b = sos[s, :3]
        a = sos[s, 3:]
        # lfilter_zi for a biquad: solve (I - A^T) zi = B
        IminusA = np.array([[1.0 + a[1], -1.0], [a[2], 1.0]])
        B = np.array([b[1] - a[1] * b[0], b[2] - a[2] * b[0]])
        zi[s] = scale * np.linalg.solve(IminusA, B)
        scale *= b.sum() / a.sum()

    def run(sig):
        sig = sig.copy()
        flat = sig.reshape(-1, sig.shape[-1])
        for row in flat:
            z = zi * row[0]
            for n in range(row.shape[0]):
                v = row[n]
                for s in range(n_sections):
                    b0, b1, b2, _, a1, a2 = sos[s]
                    out = b0 * v + z[s, 0]
                    z[s, 0] = b1 * v - a1 * out + z[s, 1]
                    z[s, 1] = b2 * v - a2 * out
                    v = out
                row[n] = v
        return sig

    fwd = run(ext)
    bwd = run(fwd[..., ::-1])[..., ::-1]
    return bwd[..., padlen:-padlen]


# ----------------------------------------------------------------------------
# applyFilter   (script/mfcc.py:29-135 == script/calc.py:23-129)
# ----------------------------------------------------------------------------


def applyFilter(x, sr, /, *, filt="iir", cutOff=[None], filtLen=6, filtType="low", polyOrd=3, coeffs=None):
    """Restatement of ``applyFilter``; same validation order and messages."""
    if (filt is None) | (cutOff is None) | (cutOff is None):
        if cutOff is None:
            raise Exception("Cannot apply filter without specifying a cut Off freq. (CutOff is None).")
        else:
            raise Exception(
                "Cannot apply filter without specifying a filter method among iir, fir and  sg (filt is None)."
            )
    filtTypes = np.array(["bandpass", "lowpass", "highpass"])
    try:
        filtType = filtTypes[np.argwhere([t.startswith(filtType) for t in filtTypes]).flatten()][0]
    except Exception:
        raise Exception("filtType must be one among: lowpass, highpass, bandpass. Partial matches allowed.")
    if any((sr / 2) <= np.array(cutOff)):
        raise Exception(
            "Cut off frequencies must be smaller than the half of the sampling freq. of the signal submitted to the filter"
        )
    if (len(cutOff) > 0) & (any(np.diff(cutOff) <= 0)):
        raise Exception("If two cut off freqs are provided: cutOff[0]<cutOff[1]")
    cutOff = np.array(cutOff)
    y = None
    ok = ((len(cutOff) == 1) and ((filtType == "lowpass") | (filtType == "highpass"))) | (
        (len(cutOff) == 2) and (filtType == "bandpass")
    )
    if filt == "iir":
        if coeffs is None:
            w = cutOff / (sr / 2)
            if ok:
                sos = scipy.signal.butter(filtLen, w, btype=filtType, output="sos")
            else:
                raise Exception(
                    "only one or two cut off frequencies allowed. If two freqs are provided, filtType must be bandpass"
                )
        y = scipy.signal.sosfiltfilt(sos, x)  # NameError if coeffs given: reference quirk (mfcc.py:100-111)
    if filt == "fir":
        if coeffs is None:
            w = cutOff / (sr / 2)
            if ok:
                bFil = scipy.signal.firwin(filtLen, w, window=("kaiser", 7.4), pass_zero=filtType)
            else:
                raise Exception(
                    "only one or two cut off frequencies allowed. If two freqs are provided, filtType must be bandpass"
                )
        y = scipy.signal.filtfilt(bFil, 1, x)
    if filt == "sg":
        if len(cutOff) == 1:
            y = scipy.signal.savgol_filter(x, filtLen, polyOrd, deriv=0, mode="interp")
        else:
            raise Exception("sg (savitsky Golay) filters can only be lowpass (one cutOff freq allowed)")
    return y


# ----------------------------------------------------------------------------
# get_MFCCS_change   (script/mfcc.py:291-427)
# ----------------------------------------------------------------------------


def get_MFCCS_change(
    audioIn,
    sigSr,
    /,
    *,
    channelN=0,
    tStep=0.001,
    winLen=0.025,
    n_mfcc=13,
    n_fft=512,
    minFreq=100,
    maxFreq=10000,
    removeFirst=1,
    filtCutoff=12,
    filtOrd=6,
    diffMethod="grad",
    outFilter="iir",
    outFiltType="low",
    outFiltCutOff=[None],
    outFiltLen=6,
    outFiltPolyOrd=3,
    n_mels=128,
    return_features=False,
):
    """Restatement of ``get_MFCCS_change`` for array input (file decoding,
    script/mfcc.py:372-373, is outside the measured path).

    ``n_mels`` is an additive keyword (the reference never passes it, so
    librosa's default 128 applies); ``return_features`` additionally returns the
    intermediates for stage-by-stage parity tests.
    """
    if isinstance(audioIn, str):
        raise TypeError("the oracle takes arrays; file decode is outside the path")
    myAudio = np.asarray(audioIn)
    if len(np.shape(myAudio)) > 1:  # mfcc.py:377-380
        y = myAudio[channelN, :]
    else:
        y = myAudio
    win_length, hop_length = frame_sizes(sigSr, winLen, tStep)  # mfcc.py:382-384
    myMfccs, inter = mfcc(
        y,
        sigSr,
        n_mfcc=n_mfcc,
        win_length=win_length,
        hop_length=hop_length,
        n_fft=n_fft,
        fmin=minFreq,
        fmax=maxFreq,
        n_mels=n_mels,
        return_intermediates=True,
    )  # mfcc.py:387
    full_mfcc = myMfccs
    T = np.round(np.multiply(np.arange(1, np.shape(myMfccs)[1] + 1), tStep) + winLen / 2, 4)  # mfcc.py:390
    if removeFirst:  # mfcc.py:393-395
        myMfccs = myMfccs[1:, :]
    cutOffNorm = filtCutoff / ((1 / tStep) / 2)  # mfcc.py:398
    sos = scipy.signal.butter(filtOrd, cutOffNorm, btype="low", output="sos")  # mfcc.py:400
    filtMffcs = scipy.signal.sosfiltfilt(sos, myMfccs)  # mfcc.py:402
    if diffMethod == "grad":  # mfcc.py:405-412
        myDiff = np.gradient(filtMffcs, axis=1)
    else:
        myDiff = scipy.signal.savgol_filter(filtMffcs, 3, 2, deriv=1, axis=1, mode="interp")
    totChange = np.sqrt(np.sum(myDiff**2, 0)) / np.shape(myMfccs)[0]  # mfcc.py:415
    rawChange = totChange
    if outFilter is None:  # mfcc.py:417-425
        totChange = scipy.signal.sosfiltfilt(sos, totChange)
    else:
        totChange = applyFilter(
            totChange,
            1 / tStep,
            filt=outFilter,
            filtType=outFiltType,
            cutOff=outFiltCutOff,
            filtLen=outFiltLen,
            polyOrd=outFiltPolyOrd,
        )
    if return_features:
        feats = dict(inter)
        feats.update(mfcc=full_mfcc, filt_mfcc=filtMffcs, diff=myDiff, raw_change=rawChange, sos=sos)
        return totChange, T, feats
    return totChange, T


# ----------------------------------------------------------------------------
# get_velocity   (script/calc.py:593-650)
# ----------------------------------------------------------------------------


def _fd_weights(offsets: np.ndarray, deriv: int) -> np.ndarray:
    """Finite-difference weights on integer ``offsets`` for the ``deriv``-th
    derivative (unit spacing): solve the Taylor/Vandermonde system, as findiff's
    ``coefficients`` does."""
    offsets = np.asarray(offsets, dtype=np.float64)
    n = len(offsets)
    A = np.vander(offsets, n, increasing=True).T
    rhs = np.zeros(n)
    rhs[deriv] = math.factorial(deriv)
    return np.linalg.solve(A, rhs)


def findiff_stencils(deriv: int, acc: int):
    """(center, forward, backward) offset/weight pairs of ``findiff.coefficients``.

    center: offsets -p..p with p = (deriv+1)//2 - 1 + acc//2;
    forward: offsets 0..(2p + extra) with the same formal accuracy, where findiff
    uses ``num_coef = 2*p + 1`` central points and ``num_coef (+1 if deriv even)``
    one-sided points.
    """
    p = (deriv + 1) // 2 - 1 + acc // 2
    c_off = np.arange(-p, p + 1)
    num_coef = 2 * p + 1
    if deriv % 2 == 0:
        num_coef += 1
    f_off = np.arange(0, num_coef)
    b_off = -f_off[::-1]
    return (
        (c_off, _fd_weights(c_off, deriv)),
        (f_off, _fd_weights(f_off, deriv)),
        (b_off, _fd_weights(b_off, deriv)),
    )


def _findiff_apply(x: np.ndarray, h: float, deriv: int, acc: int) -> np.ndarray:
    """``FinDiff(0, h, deriv, acc=acc)(x)`` along axis 0 [upstream findiff]:
    central stencil in the interior, equal-accuracy one-sided stencils on the
    first/last ``p`` points."""
    x = np.asarray(x, dtype=np.float64)
    (c_off, c_w), (f_off, f_w), (b_off, b_w) = findiff_stencils(deriv, acc)
    p = int(c_off[-1])
    n = x.shape[0]
    y = np.zeros_like(x)
    for off, w in zip(c_off, c_w):
        y[p : n - p] += w * x[p + off : n - p + off]
    for off, w in zip(f_off, f_w):
        y[:p] += w * x[off : p + off]
    for off, w in zip(b_off, b_w):
        y[n - p :] += w * x[n - p + off : n + off]
    return y / h**deriv


def get_velocity(x, sr, difference=1, method="gradient", width=3, accOrder=2, polyOrder=2):
    """Restatement of ``get_velocity`` (script/calc.py:635-650)."""
    if method == "finDiff":
        y = _findiff_apply(x, 1 / sr, difference, accOrder)
    elif method == "sg":
        y = scipy.signal.savgol_filter(x, width, polyOrder, deriv=difference, axis=0, mode="interp")
    elif method == "gradient":
        for _ in range(difference):
            x = np.gradient(x, 1 / sr)
        y = x
    else:
        raise ValueError("Méthode inconnue. Utilisez 'gradient', 'sg' ou 'finDiff'.")
    return y


# ----------------------------------------------------------------------------
# calculate_amplitude_envelope / get_amplitude
# (script/calc.py:221-343 == script/mfcc.py:137-259)
# ----------------------------------------------------------------------------


def rms_frames(x: np.ndarray, frame_length: int, hop_length: int, center: bool = True) -> np.ndarray:
    """``librosa.feature.rms(y=x, frame_length, hop_length, center, pad_mode='constant')``
    flattened: float32 ``sqrt(mean(frame**2))``."""
    x = np.asarray(x)
    if center:
        pad = frame_length // 2
        x = np.concatenate([np.zeros(pad, x.dtype), x, np.zeros(pad, x.dtype)])
    if len(x) < frame_length:
        raise ValueError("Input is too short (n=%d) for frame_length=%d" % (len(x), frame_length))
    T = 1 + (len(x) - frame_length) // hop_length
    frames = np.lib.stride_tricks.as_strided(
        x, shape=(frame_length, T), strides=(x.strides[0], x.strides[0] * hop_length), writeable=False
    )
    power = np.mean(np.square(frames, dtype=np.float32), axis=-2, keepdims=True)
    return np.sqrt(power).flatten()


def calculate_amplitude_envelope(
    x,
    sr,
    /,
    *,
    method="RMS",
    winLen=0.1,
    hopLen=0.01,
    center=True,
    outFilter=None,
    outFiltType="low",
    outFiltCutOff=[12],
    outFiltLen=6,
    outFiltPolyOrd=3,
):
    """Restatement of ``calculate_amplitude_envelope`` (methods RMS and Hilb;
    RMSpraat needs Praat and is out of scope)."""
    if method == "Hilb":
        amp = np.abs(scipy.signal.hilbert(x))
        ampT = np.arange(len(x)) / sr
        ampSr = sr
    elif method == "RMSpraat":
        raise NotImplementedError("RMSpraat calls Praat (parselmouth); out of scope")
    elif method == "RMS":
        frLen = int(hopLen * sr)
        winLen = int(winLen * sr)
        amp = rms_frames(x, winLen, frLen, center)
    if (method != "hilb") & (method != "RMSpraat"):  # sic: lower-case typo, calc.py:333
        ampT = np.arange(len(amp)) * hopLen
        ampSr = 1 / hopLen
    if outFilter is not None:
        amp = applyFilter(
            amp, ampSr, filt=outFilter, filtType=outFiltType, cutOff=outFiltCutOff, filtLen=outFiltLen, polyOrd=outFiltPolyOrd
        )
    return amp, ampT


get_amplitude = calculate_amplitude_envelope  # script/mfcc.py:137-259 has the identical body


# ----------------------------------------------------------------------------
# Appendix B  modulation spectrum of the MFCC trajectories (EXTENSION: no
# counterpart in the reference, parity unpinned by construction)
# ----------------------------------------------------------------------------

MODULATION_BANDS_HZ = ((0.5, 2.0), (2.0, 4.0), (4.0, 8.0), (8.0, 16.0), (16.0, 32.0))


def modspec_sizes(T: int, frame_rate: float, mod_win_s: float = 1.0, mod_hop_s: float = 0.5):
    Lw = int(round(mod_win_s * frame_rate))
    Hw = max(1, int(round(mod_hop_s * frame_rate)))
    n_mod_fft = 1 << max(1, (Lw - 1).bit_length())
    n_win = 1 + (T - Lw) // Hw if T >= Lw else 0
    return Lw, Hw, n_mod_fft, n_win


def modulation_spectrum(
    M: np.ndarray,
    frame_rate: float,
    *,
    mod_win_s: float = 1.0,
    mod_hop_s: float = 0.5,
    bands_hz=MODULATION_BANDS_HZ,
):
    """|rfft| of mean-removed, Hann-windowed MFCC trajectory windows.

    M: ``[n_coef, T]``.  Returns ``(mag [n_coef, n_win, n_mod_fft/2+1] f32,
    band_energy [n_win, n_bands] f32, mod_freqs [n_mod_fft/2+1] f64)``; computed
    in float64 and cast.
    """
    M = np.asarray(M, dtype=np.float64)
    C, T = M.shape
    Lw, Hw, nfft, n_win = modspec_sizes(T, frame_rate, mod_win_s, mod_hop_s)
    nb = nfft // 2 + 1
    freqs = np.arange(nb) * frame_rate / nfft
    if n_win <= 0:
        return np.zeros((C, 0, nb), np.float32), np.zeros((0, len(bands_hz)), np.float32), freqs
    w = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(Lw) / Lw)
    idx = (np.arange(n_win) * Hw)[:, None] + np.arange(Lw)[None, :]
    seg = M[:, idx]  # [C, n_win, Lw]
    seg = (seg - seg.mean(axis=-1, keepdims=True)) * w
    X = np.fft.rfft(seg, n=nfft, axis=-1)
    mag = np.abs(X)
    p = mag**2
    E = np.zeros((n_win, len(bands_hz)))
    for b, (lo, hi) in enumerate(bands_hz):
        sel = (freqs >= lo) & (freqs < hi)
        E[:, b] = p[:, :, sel].sum(axis=(0, 2))
    return mag.astype(np.float32), E.astype(np.float32), freqs


# ----------------------------------------------------------------------------
# Batch feature bundle used by the parity tests and the CPU baseline
# ----------------------------------------------------------------------------


def mfcc_features(
    y: np.ndarray,
    sr: float,
    *,
    tStep: float = 0.01,
    winLen: float = 0.025,
    n_fft: int = 512,
    n_mels: int = 40,
    n_mfcc: int = 13,
    fmin: float = 0.0,
    fmax: float | None = None,
    removeFirst: int = 1,
    filtCutoff: float = 12,
    filtOrd: int = 6,
    mod_win_s: float = 1.0,
    mod_hop_s: float = 0.5,
):
    """Everything the GPU batch API produces for one clip: log-mel, MFCC, delta
    (``get_velocity(., sr=1.0)`` == ``np.gradient`` along time), MFCC-change
    (``get_MFCCS_change`` with the GUI's iir output filter at ``filtCutoff``) and
    the modulation spectrum."""
    if fmax is None:
        fmax = sr / 2
    tot, T, f = get_MFCCS_change(
        y,
        sr,
        tStep=tStep,
        winLen=winLen,
        n_mfcc=n_mfcc,
        n_fft=n_fft,
        minFreq=fmin,
        maxFreq=fmax,
        removeFirst=removeFirst,
        filtCutoff=filtCutoff,
        filtOrd=filtOrd,
        diffMethod="grad",
        outFilter="iir",
        outFiltType="low",
        outFiltCutOff=[filtCutoff],
        outFiltLen=filtOrd,
        n_mels=n_mels,
        return_features=True,
    )
    delta = np.gradient(f["mfcc"], axis=1)
    mag, E, freqs = modulation_spectrum(f["mfcc"], 1.0 / tStep, mod_win_s=mod_win_s, mod_hop_s=mod_hop_s)
    return {
        "logmel": f["logmel"],
        "melspec": f["melspec"],
        "power": f["power"],
        "mfcc": f["mfcc"],
        "delta": delta,
        "filt_mfcc": f["filt_mfcc"],
        "totChange": tot,
        "T": T,
        "modspec": mag,
        "band_energy": E,
        "mod_freqs": freqs,
    }
