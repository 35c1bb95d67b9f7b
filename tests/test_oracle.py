"""CPU tests of the oracle itself: cross-checks against two independent
librosa-compatible implementations installed in the image (torchaudio,
transformers.audio_utils), scipy, analytic known answers, and the committed
golden vectors (the reference ships none -- SURVEY.md section 4)."""

import os

import numpy as np
import pytest
import scipy.fftpack
import scipy.signal

import oracle
from modulation_mfcc_b200.synth import synth_clip

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_frame_sizes_truncate():
    # script/mfcc.py:382-384: Python int() truncation of the float product
    assert oracle.frame_sizes(44100, 0.025, 0.01) == (1102, 441)
    assert oracle.frame_sizes(22050, 0.025, 0.01) == (551, 220)
    assert oracle.frame_sizes(16000, 0.025, 0.01) == (400, 160)
    assert oracle.frame_sizes(10000, 0.025, 0.005) == (250, 50)


@pytest.mark.parametrize("n,hop", [(160000, 160), (100000, 50), (441000, 441), (1050, 50), (5, 160)])
def test_frame_count(n, hop):
    assert oracle.n_frames(n, 512, hop) == 1 + n // hop
    y = np.zeros(n, np.float32)
    assert oracle.stft_power(y, 512, hop, 400).shape == (257, 1 + n // hop)


def test_window_padding():
    w = oracle.padded_hann(400, 512)
    assert w.shape == (512,) and np.all(w[:56] == 0) and np.all(w[456:] == 0) and w[56] == 0.0
    assert np.allclose(w[56:456], scipy.signal.get_window("hann", 400, fftbins=True))
    with pytest.raises(ValueError):
        oracle.padded_hann(600, 512)


def test_mel_matches_torchaudio_and_transformers():
    import torch
    import torchaudio
    from transformers.audio_utils import mel_filter_bank

    for sr, n_fft, n_mels, fmin, fmax in [(16000, 512, 40, 0.0, 8000.0), (44100, 2048, 128, 0.0, 22050.0), (10000, 512, 128, 100.0, 5000.0)]:
        W = oracle.mel_filterbank(sr, n_fft, n_mels, fmin, fmax)
        ta = torchaudio.functional.melscale_fbanks(n_fft // 2 + 1, fmin, fmax, n_mels, sr, norm="slaney", mel_scale="slaney").T.numpy()
        hf = mel_filter_bank(n_fft // 2 + 1, n_mels, fmin, fmax, sr, norm="slaney", mel_scale="slaney").T
        assert np.max(np.abs(W - ta)) < 1e-6
        assert np.max(np.abs(W - hf)) < 1e-6


def test_mel_fmax_above_nyquist_leaves_empty_filters():
    W = oracle.mel_filterbank(10000, 512, 128, 100, 10000)
    assert int((W.sum(axis=1) == 0).sum()) == 26  # 26 of 128 filters are all-zero
    W = oracle.mel_filterbank(16000, 512, 128, 100, 10000)
    assert int((W.sum(axis=1) == 0).sum()) == 7


def test_dct_matrix_matches_scipy():
    x = np.random.default_rng(0).standard_normal((40, 7))
    D = oracle.dct_ortho_matrix(13, 40)
    assert np.max(np.abs(D @ x - scipy.fftpack.dct(x, axis=0, type=2, norm="ortho")[:13])) < 1e-12


def test_mfcc_matches_torchaudio():
    import torch
    import torchaudio

    y = synth_clip(0, 32000, 16000)
    M, inter = oracle.mfcc(y, 16000, n_mfcc=13, win_length=400, hop_length=160, n_fft=512, fmin=0, fmax=8000, n_mels=40, return_intermediates=True)
    ms = torchaudio.transforms.MelSpectrogram(16000, n_fft=512, win_length=400, hop_length=160, f_min=0.0, f_max=8000.0, n_mels=40,
                                              center=True, pad_mode="constant", power=2.0, norm="slaney", mel_scale="slaney")
    S = ms(torch.from_numpy(y))
    db = torchaudio.transforms.AmplitudeToDB("power", top_db=80.0)(S)
    Mt = (db.T @ torchaudio.functional.create_dct(13, 40, "ortho")).T.numpy()
    assert np.max(np.abs(S.numpy() - inter["melspec"]) / inter["melspec"]) < 2e-4  # torch's fp32 FFT noise
    assert np.max(np.abs(Mt - M)) < 5e-4


def test_power_to_db_clamp():
    S = np.array([[1.0, 1e-3], [1e-12, 1e-9]], np.float32)
    db = oracle.power_to_db(S)
    assert np.allclose(db, [[0.0, -30.0], [-80.0, -80.0]], atol=1e-5)  # amin floor -100 then clamp to max-80
    assert np.allclose(oracle.power_to_db(np.zeros((3, 4), np.float32)), -100.0)


def test_pure_tone_peaks_at_its_bin():
    sr, n = 16000, 16000
    t = np.arange(n) / sr
    y = np.sin(2 * np.pi * (sr / 512 * 40) * t).astype(np.float32)
    P = oracle.stft_power(y, 512, 160, 400)
    assert np.all(np.argmax(P[:, 5:-5], axis=0) == 40)


def test_sosfiltfilt_restated_and_short_input():
    rng = np.random.default_rng(1)
    x = rng.standard_normal((3, 200))
    for sos in (scipy.signal.butter(6, 0.24, output="sos"), scipy.signal.butter(3, [0.1, 0.4], btype="band", output="sos")):
        assert np.max(np.abs(oracle.sosfiltfilt_restated(sos, x) - scipy.signal.sosfiltfilt(sos, x))) < 1e-11
    sos = scipy.signal.butter(6, 0.24, output="sos")
    with pytest.raises(ValueError, match="padlen, which is 21"):
        scipy.signal.sosfiltfilt(sos, x[:, :21])
    with pytest.raises(ValueError, match="padlen, which is 21"):
        oracle.sosfiltfilt_restated(sos, x[:, :21])
    oracle.sosfiltfilt_restated(sos, x[:, :22])


def test_get_MFCCS_change_shapes_and_anchors():
    y = synth_clip(1, 20000, 10000)
    kw = dict(tStep=0.005, winLen=0.025, n_mfcc=13, n_fft=512, minFreq=100, maxFreq=10000, outFiltCutOff=[12])
    tot, T = oracle.get_MFCCS_change(y, 10000, **kw)
    assert tot.shape == T.shape == (401,) and tot.dtype == np.float64
    assert T[0] == 0.0175 and T[1] == 0.0225  # round(k*tStep + winLen/2, 4), k from 1 (script/mfcc.py:390)
    assert np.all(tot >= -1e-9)
    with pytest.raises(TypeError):  # signature default outFiltCutOff=[None] (script/mfcc.py:308,:93)
        oracle.get_MFCCS_change(y, 10000, tStep=0.005)


def test_applyFilter_validation_messages():
    x = np.random.default_rng(2).standard_normal(100)
    with pytest.raises(Exception, match="CutOff is None"):
        oracle.applyFilter(x, 100.0, cutOff=None)
    with pytest.raises(Exception, match="filt is None"):
        oracle.applyFilter(x, 100.0, filt=None, cutOff=[5])
    with pytest.raises(Exception, match="filtType must be one among"):
        oracle.applyFilter(x, 100.0, cutOff=[5], filtType="notch")
    with pytest.raises(Exception, match="smaller than the half"):
        oracle.applyFilter(x, 100.0, cutOff=[50])
    with pytest.raises(Exception, match=r"cutOff\[0\]<cutOff\[1\]"):
        oracle.applyFilter(x, 100.0, cutOff=[20, 10], filtType="band")
    with pytest.raises(Exception, match="only one or two cut off"):
        oracle.applyFilter(x, 100.0, cutOff=[10, 20], filtType="low")
    assert oracle.applyFilter(x, 100.0, cutOff=[5], filtType="lo").shape == x.shape  # prefix match


def test_findiff_stencils_known_values():
    (co, cw), (fo, fw), (bo, bw) = oracle.findiff_stencils(1, 2)
    assert np.allclose(cw, [-0.5, 0, 0.5]) and np.allclose(fw, [-1.5, 2, -0.5]) and np.allclose(bw, [0.5, -2, 1.5])
    (co, cw), (fo, fw), (bo, bw) = oracle.findiff_stencils(2, 2)
    assert np.allclose(cw, [1, -2, 1]) and np.allclose(fw, [2, -5, 4, -1]) and np.allclose(bw, [-1, 4, -5, 2])
    x = np.linspace(0, 1, 50) ** 3
    assert np.allclose(oracle.get_velocity(x, 49.0, 1, "finDiff", accOrder=4), 3 * np.linspace(0, 1, 50) ** 2, atol=1e-9)
    with pytest.raises(ValueError, match="Méthode inconnue"):
        oracle.get_velocity(x, 1.0, method="x")


def test_rms_envelope_known_answer():
    y = np.ones(1600, np.float32) * 0.5
    amp, t = oracle.calculate_amplitude_envelope(y, 16000.0, winLen=0.01, hopLen=0.005, center=False)
    assert np.allclose(amp, 0.5) and amp.dtype == np.float32 and np.allclose(np.diff(t), 0.005)
    amp, _ = oracle.calculate_amplitude_envelope(y, 16000.0, winLen=0.01, hopLen=0.005, center=True)
    assert np.isclose(amp[0], 0.5 * np.sqrt(0.5), rtol=1e-6)  # half of the first frame is padding


def test_modulation_spectrum_known_answer():
    fr, T = 100.0, 1001
    t = np.arange(T) / fr
    M = np.stack([3.0 + 2.0 * np.sin(2 * np.pi * 5.0 * t), np.zeros(T)])
    mag, E, freqs = oracle.modulation_spectrum(M, fr)
    assert mag.shape == (2, 19, 65) and E.shape == (19, 5)
    k = np.argmax(mag[0, 3])
    assert abs(freqs[k] - 5.0) < fr / 128
    assert np.all(mag[1] == 0)
    assert np.argmax(E[3]) == 2  # the [4, 8) Hz band
    # mean removal: a constant trajectory has no modulation energy
    assert np.max(oracle.modulation_spectrum(np.full((1, T), 7.0), fr)[0]) < 1e-9


@pytest.mark.parametrize("name", ["cfg1_16k_40mel", "cfg3_44k_128mel", "gui_default_10k", "cfg4_long_hop"])
def test_golden_vectors(name):
    import sys
    sys.path.insert(0, GOLDEN)
    from make_golden import CASES

    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    seed, sr, secs, kw = CASES[name]
    y = synth_clip(seed, int(sr * secs), sr)
    assert np.array_equal(y[:64], g["y_head"])
    f = oracle.mfcc_features(y, sr, **kw)
    step = 10 if name == "cfg4_long_hop" else 1
    assert np.allclose(f["logmel"], g["logmel"], atol=2e-4)
    assert np.allclose(f["mfcc"], g["mfcc"], atol=2e-4)
    assert np.allclose(f["totChange"], g["totChange"], atol=1e-6)
    assert np.array_equal(f["T"], g["T"])
    assert np.allclose(f["modspec"][:, ::step], g["modspec"], atol=2e-3)
    assert np.allclose(f["band_energy"], g["band_energy"], rtol=1e-4, atol=1e-2)
