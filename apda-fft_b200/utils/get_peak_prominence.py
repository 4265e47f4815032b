"""Drop-in for the reference module ``utils/get_peak_prominence.py`` (flexible structures).

Same public names and call signatures; the work runs in libapda_b200.so (kernel K3, prominence picker).
"""
from __future__ import annotations

import ctypes

import numpy as np

from apda_fft_b200 import _cabi
from apda_fft_b200.records import prominence_dicts, record_dtype
from metrics.fft_iterativa import pack_spectrum

_p = ctypes.c_void_p


def _mags(magnitudes):
    return np.ascontiguousarray(magnitudes, dtype=np.float64)


def calculate_prominence(magnitudes, peak_idx):
    """reference :32-54."""
    m = _mags(magnitudes)
    idx = peak_idx if peak_idx >= 0 else peak_idx + len(m)
    out = ctypes.c_double()
    _cabi.default_context().call("apda_prominence_f64_host", _p(m.ctypes.data), m.shape[0], idx, ctypes.byref(out))
    return out.value


def calculate_half_power_width_prominenceBased(magnitudes, prominence, peak_idx, fs, n):
    """reference :89-112 - half-power width in Hz of the peak measured from its valley."""
    m = _mags(magnitudes)
    bins = ctypes.c_int64()
    _cabi.default_context().call("apda_half_power_bins_f64_host", _p(m.ctypes.data), m.shape[0], float(prominence),
                                 peak_idx, ctypes.byref(bins))
    return bins.value * (fs / n)


def get_top_peaks_prominence(res_fft, fs, k=4):
    """reference :149-226 - up to k peaks, descending rounded magnitude, each
    {freq, mag, prominence, damping, q-factor, idx}."""
    n = len(res_fft)
    z = pack_spectrum(res_fft)
    want = max(int(k), 1)   # the reference tests len(final_peaks) >= k only after appending
    # any k, as in the reference: a window of n bins cannot hold more than n/8 + 8 peaks (include/apda_b200.h
    # APDA_MAX_PEAKS), so a larger k asks for "all of them" and the record is sized for that
    want = min(want, _cabi.max_peaks(n))
    cap = max(5, want)
    rec = np.zeros(1, dtype=record_dtype(cap))
    _cabi.default_context().call("apda_peaks_prominence_f64_host", _p(z.ctypes.data), n, 1, float(fs), _p(0), want, cap,
                                 _p(rec.ctypes.data))
    _cabi.check_record_status(rec)
    return prominence_dicts(rec[0], fs, n)
