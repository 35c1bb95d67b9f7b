"""Case table shared by the reference-pinned tests.

Each case names a function of the reference's ``script/mfcc.py`` / ``script/calc.py``, a seeded input and the
keyword arguments.  ``tests/golden/make_ref_golden.py`` runs the UNMODIFIED reference modules on these cases
(``oracle/ref_loader.py``) and commits the outputs as ``tests/golden/ref_outputs.npz``; the CPU tests hold the
oracle to those outputs bit for bit, the GPU tests hold the CUDA path to them within north_star's tolerances.
"""

from __future__ import annotations

import numpy as np

KW_GUI = dict(channelN=0, tStep=0.005, winLen=0.025, n_mfcc=13, n_fft=512, minFreq=100, maxFreq=10000, removeFirst=1,
              filtCutoff=12, filtOrd=6, diffMethod="grad", outFilter="iir", outFiltType="low", outFiltCutOff=[12],
              outFiltLen=6, outFiltPolyOrd=3)


def _clip(seed, n, sr):
    from modulation_mfcc_b200.synth import synth_clip

    return synth_clip(seed, n, sr)


def _walk(seed, n, cols=None):
    rng = np.random.default_rng(seed)
    shape = (n,) if cols is None else (n, cols)
    return rng.standard_normal(shape).cumsum(axis=0)


def _kw(**over):
    kw = dict(KW_GUI)
    kw.update(over)
    return kw


# name -> (module, function, input thunk, positional tail, kwargs)
CASES = {}


def _add(name, module, func, make_x, args, kwargs):
    assert name not in CASES
    CASES[name] = (module, func, make_x, args, kwargs)


# ---- get_MFCCS_change (script/mfcc.py:291-427): the GUI call of main.py:750-769 and every branch after it
_add("change_gui", "mfcc", "get_MFCCS_change", lambda: _clip(1, 40000, 10000), (10000,), _kw())
for _i, _over in enumerate([
    dict(outFilter=None),
    dict(diffMethod="sg"),
    dict(removeFirst=0),
    dict(outFilter="fir", outFiltLen=11, outFiltCutOff=[12]),
    dict(outFilter="sg", outFiltLen=7, outFiltPolyOrd=3, outFiltCutOff=[12]),
    dict(outFilter="iir", outFiltType="band", outFiltCutOff=[2, 20], outFiltLen=4),
    dict(outFilter="iir", outFiltType="high", outFiltCutOff=[5], outFiltLen=3),
    dict(tStep=0.01, n_fft=1024, winLen=0.04),
    dict(filtCutoff=8, filtOrd=4, outFiltCutOff=[20], outFiltLen=2),
    dict(n_mfcc=20, minFreq=0, maxFreq=5000),
]):
    _add(f"change_var{_i}", "mfcc", "get_MFCCS_change", lambda: _clip(2, 30000, 10000), (10000,), _kw(**_over))
_add("change_multichannel", "mfcc", "get_MFCCS_change",
     lambda: np.stack([_clip(3, 20000, 10000), _clip(4, 20000, 10000)]), (10000,), _kw(channelN=1))
_add("change_t22", "mfcc", "get_MFCCS_change", lambda: _clip(9, 21 * 50, 10000), (10000,), _kw())
_add("change_16k", "mfcc", "get_MFCCS_change", lambda: _clip(0, 32000, 16000), (16000,),
     _kw(tStep=0.01, minFreq=0, maxFreq=8000))
_add("change_44k", "mfcc", "get_MFCCS_change", lambda: _clip(5, 44100, 44100), (44100,),
     _kw(tStep=0.01, n_fft=2048, n_mfcc=20, minFreq=0, maxFreq=22050))

# ---- applyFilter (script/mfcc.py:29-135 and its duplicate script/calc.py:23-129)
for _i, _k in enumerate([
    dict(filt="iir", cutOff=[12], filtLen=6),
    dict(filt="fir", cutOff=[12], filtLen=21),
    dict(filt="fir", cutOff=[5, 30], filtLen=31, filtType="band"),
    dict(filt="sg", cutOff=[12], filtLen=9, polyOrd=3),
    dict(filt="iir", cutOff=[20], filtLen=4, filtType="high"),
    dict(filt="iir", cutOff=[3, 40], filtLen=3, filtType="bandp"),
    dict(filt="fir", cutOff=[25], filtLen=15, filtType="h"),
]):
    for _m in ("mfcc", "calc"):
        _add(f"filter{_i}_{_m}", _m, "applyFilter", lambda: _walk(5, 777), (200.0,), _k)

# ---- get_velocity (script/calc.py:593-650); the GUI passes sr=1.0 (main.py:680-688, 704-712)
for _i, _k in enumerate([
    dict(method="gradient", difference=1), dict(method="gradient", difference=2),
    dict(method="sg", difference=1, width=5, polyOrder=2), dict(method="sg", difference=2, width=7, polyOrder=3),
    dict(method="finDiff", difference=1, accOrder=2), dict(method="finDiff", difference=2, accOrder=2),
    dict(method="finDiff", difference=1, accOrder=4), dict(method="finDiff", difference=2, accOrder=4),
]):
    for _sr in (1.0, 200.0):
        _add(f"velocity{_i}_sr{int(_sr)}", "calc", "get_velocity", lambda: _walk(5, 777), (_sr,), _k)
# axis-0 behaviour on 2-D input (calc.py:639 axis=0; FinDiff(0, ...))
_add("velocity_2d_sg", "calc", "get_velocity", lambda: _walk(6, 300, 4), (1.0,), dict(method="sg", difference=1, width=5, polyOrder=2))
_add("velocity_2d_fd", "calc", "get_velocity", lambda: _walk(6, 300, 4), (50.0,), dict(method="finDiff", difference=1, accOrder=2))

# ---- amplitude envelope (script/calc.py:221-343, script/mfcc.py:137-259)
_add("env_rms_calc", "calc", "calculate_amplitude_envelope", lambda: _clip(6, 48000, 16000), (16000,), dict())
_add("env_rms_mfcc", "mfcc", "get_amplitude", lambda: _clip(6, 48000, 16000), (16000,), dict())
_add("env_rms_filt", "calc", "calculate_amplitude_envelope", lambda: _clip(7, 48000, 16000), (16000,),
     dict(winLen=0.05, hopLen=0.005, center=False, outFilter="iir", outFiltCutOff=[12]))
_add("env_rms_fir", "calc", "calculate_amplitude_envelope", lambda: _clip(7, 30000, 10000), (10000,),
     dict(outFilter="fir", outFiltCutOff=[10], outFiltLen=15))
_add("env_hilb", "calc", "calculate_amplitude_envelope", lambda: _clip(8, 12345, 16000), (16000,), dict(method="Hilb"))
_add("env_hilb_even", "calc", "calculate_amplitude_envelope", lambda: _clip(8, 16000, 16000), (16000,), dict(method="Hilb"))

# ---- error behaviour: (case name, module, function, input thunk, args, kwargs, exception type, message fragment)
ERROR_CASES = [
    ("change_default_cutoff", "mfcc", "get_MFCCS_change", lambda: _clip(2, 30000, 10000), (10000,), _kw(outFiltCutOff=[None]), TypeError, ""),
    ("change_cutoff_high", "mfcc", "get_MFCCS_change", lambda: _clip(2, 30000, 10000), (10000,), _kw(outFiltCutOff=[200]), Exception, "Cut off frequencies must be smaller"),
    ("change_bad_type", "mfcc", "get_MFCCS_change", lambda: _clip(2, 30000, 10000), (10000,), _kw(outFiltType="notch"), Exception, "filtType must be one among"),
    ("change_t21", "mfcc", "get_MFCCS_change", lambda: _clip(2, 50 * 20, 10000), (10000,), _kw(), ValueError, "greater than padlen"),
    ("filter_two_cutoffs_low", "calc", "applyFilter", lambda: _walk(5, 300), (200.0,), dict(filt="iir", cutOff=[5, 30], filtType="low"), Exception, ""),
    ("filter_one_cutoff_band", "calc", "applyFilter", lambda: _walk(5, 300), (200.0,), dict(filt="iir", cutOff=[5], filtType="band"), Exception, ""),
    ("velocity_unknown", "calc", "get_velocity", lambda: _walk(5, 300), (1.0,), dict(method="nope"), ValueError, "Méthode inconnue"),
]


def run_case(modules, name):
    """Evaluate case ``name`` with ``modules = {'mfcc': <module>, 'calc': <module>}``; returns a tuple of arrays."""
    module, func, make_x, args, kwargs = CASES[name]
    out = getattr(modules[module], func)(make_x(), *args, **kwargs)
    if not isinstance(out, tuple):
        out = (out,)
    return tuple(np.asarray(o) for o in out)
