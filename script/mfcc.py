"""Drop-in replacement for the reference's ``script/mfcc.py``.

Same module name and public functions (``load_channel``, ``get_MFCCS_change``,
``applyFilter``, ``get_amplitude``) with the reference's signatures
(script/mfcc.py:29-39, 137-150, 262-264, 291-311), so ``script/main.py:29``
(``from mfcc import load_channel, get_MFCCS_change``) keeps working unchanged.
The arithmetic runs on a B200 through ``modulation_mfcc_b200``; librosa,
parselmouth and pyqtgraph are no longer imported here.
"""

import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from modulation_mfcc_b200.api import (  # noqa: E402,F401
    applyFilter,
    get_amplitude,
    get_MFCCS_change,
    get_MFCCS_change_batch,
    load_channel,
    mfcc_features_batch,
)
