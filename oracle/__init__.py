"""CPU oracle for the MFCC / cepstral-modulation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the shipped
product path: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and
there only as the checker or as the timed CPU baseline.  The product package
(``modulation_mfcc_b200``) never imports this package and raises loudly when its
CUDA library is missing.

PARITY: pinned to the reference's executed code for everything but the librosa call.
The reference repository (aaron-randreth/modulation-mfcc) has no tests, golden vectors
or fixtures for this path, and ``script/mfcc.py:8,18-23,27`` import librosa, parselmouth
and pyqtgraph, none of which is installed (no network).  ``oracle/ref_loader.py`` runs
``script/mfcc.py`` and ``script/calc.py`` UNMODIFIED with stub modules for those
packages; ``tests/test_reference_pin.py`` checks this restatement bit for bit against
their outputs (``tests/golden/ref_outputs.npz``).  The one step that remains a
restatement is ``librosa.feature.mfcc`` itself (an un-vendored, un-pinned dependency,
``requirements.txt:1-12``), cross-checked against torchaudio and transformers.  This oracle

* restates the librosa call chain behind ``librosa.feature.mfcc`` /
  ``librosa.feature.rms`` (librosa >= 0.10 semantics: ``center=True``,
  ``pad_mode='constant'``, Slaney mel scale and area normalisation,
  ``power_to_db(top_db=80)``, ortho DCT-II) from its published algorithm,
* calls the *installed* scipy for the steps the reference itself delegates to
  scipy (``butter``, ``sosfiltfilt``, ``filtfilt``, ``firwin``,
  ``savgol_filter``, ``hilbert``, ``fftpack.dct``) -- scipy is the actual
  upstream implementation of those steps,
* restates findiff's finite-difference stencils for ``get_velocity``,
* follows ``script/mfcc.py:372-427`` and ``script/calc.py:593-650`` line by line
  for everything around those calls,

and is cross-checked in ``tests/test_oracle.py`` against two independent
librosa-compatible implementations that *are* installed (torchaudio 2.11
``MelSpectrogram``/``AmplitudeToDB``/``create_dct`` and
``transformers.audio_utils``) plus analytic known-answer cases.  The golden
vectors in ``tests/golden/`` were minted from this oracle by
``tests/golden/make_golden.py``.
"""

from .mfcc_oracle import (  # noqa: F401
    frame_sizes,
    padded_hann,
    stft_power,
    mel_filterbank,
    power_to_db,
    dct_ortho_matrix,
    mfcc,
    get_MFCCS_change,
    applyFilter,
    get_velocity,
    calculate_amplitude_envelope,
    get_amplitude,
    modulation_spectrum,
    mfcc_features,
    sosfiltfilt_restated,
    findiff_stencils,
    rms_frames,
    modspec_sizes,
    n_frames,
    hz_to_mel,
    mel_to_hz,
    MODULATION_BANDS_HZ,
)
