"""Drop-in replacement for the hot-path part of the reference's ``script/calc.py``.

``applyFilter`` (calc.py:23-129), ``calculate_amplitude_envelope`` (calc.py:221-343)
and ``get_velocity`` (calc.py:593-650) run on a B200 through
``modulation_mfcc_b200``.  The Praat/EMA helpers that ``script/main.py:30-36`` and
``script/ui.py:6`` also import from this module are outside the hot path
(SURVEY.md section 2, rows 6-9): ``MinMaxFinder`` is a few lines of host-side peak
picking and is provided; ``calc_formants``, ``get_f0`` and ``read_AG50x`` need
Praat / AG50x files and raise ``NotImplementedError`` here so that the import
succeeds and the failure is explicit.
"""

import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from modulation_mfcc_b200.api import (  # noqa: E402,F401
    applyFilter,
    calculate_amplitude_envelope,
    get_velocity,
)


def _out_of_scope(name):
    def f(*a, **k):
        raise NotImplementedError(
            f"calc.{name} wraps Praat/AG50x I/O and is outside the B200 hot path; use the reference's implementation"
        )

    f.__name__ = name
    return f


calc_formants = _out_of_scope("calc_formants")
get_f0 = _out_of_scope("get_f0")
read_AG50x = _out_of_scope("read_AG50x")
interp_NAN = _out_of_scope("interp_NAN")


class MinMaxFinder:
    """Local extrema of a curve inside an interval (GUI analysis on a few thousand
    points; host-side, same behaviour as the reference's class of this name)."""

    def find_in_interval(self, times, values, interval):
        start, end = interval
        t = np.asarray(times)
        v = np.asarray(values)
        keep = (start <= t) & (t <= end)
        return t[keep], v[keep]

    @staticmethod
    def _peaks(v):
        v = np.asarray(v, dtype=float)
        if len(v) < 3:
            return np.zeros(0, dtype=int)
        from scipy.signal import find_peaks

        return find_peaks(v)[0]

    def analyse_minimum(self, x, y, interval):
        if interval is None:
            print("No interval specified.")
            return [], []
        t, v = self.find_in_interval(x, y, interval)
        idx = self._peaks(-v)
        if len(idx) == 0:
            return [], []
        return t[idx], v[idx]

    def analyse_maximum(self, x, y, interval):
        if interval is None:
            print("No interval specified.")
            return [], []
        t, v = self.find_in_interval(x, y, interval)
        idx = self._peaks(v)
        if len(idx) == 0:
            return [], []
        return t[idx], v[idx]
