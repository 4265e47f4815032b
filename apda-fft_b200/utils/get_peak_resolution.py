"""Drop-in for the reference module ``utils/get_peak_resolution.py`` (rigid structures).

Same public names and call signatures; the work runs in libapda_b200.so (kernel K3, resolution picker).
"""
from __future__ import annotations

import ctypes

import numpy as np

from apda_fft_b200 import _cabi
from apda_fft_b200.records import record_dtype, resolution_dicts
from metrics.fft_iterativa import pack_spectrum

_p = ctypes.c_void_p


def width_half_magnitude(magnitudes, peak_idx):
    """reference :30-44 - run length (bins) above 0.707 * magnitudes[peak_idx]."""
    m = np.ascontiguousarray(magnitudes, dtype=np.float64)
    bins = ctypes.c_int64()
    _cabi.default_context().call("apda_half_height_bins_f64_host", _p(m.ctypes.data), m.shape[0], peak_idx,
                                 ctypes.byref(bins))
    return bins.value


def resolution(magnitudes, idx1, idx2):
    """reference :48-62 - 1.18 * |idx2 - idx1| / (w1 + w2), or 0 when both widths are 0."""
    total = width_half_magnitude(magnitudes, idx1) + width_half_magnitude(magnitudes, idx2)
    if total == 0:
        return 0
    return 1.18 * abs(idx2 - idx1) / total


def get_top_peaks_resolution(fft_res, fs, k=5):
    """reference :80-128 - up to k peaks in discovery order, each {freq, mag, idx}."""
    n = len(fft_res)
    z = pack_spectrum(fft_res)
    if k < 1:
        # the reference never enters its loop; it still evaluates the statistics first
        _cabi.check(_cabi.ERR_STATS_MEAN if n // 2 < 1 else _cabi.ERR_STATS_STDEV if n // 2 < 2 else _cabi.OK)
        return []
    want = min(int(k), _cabi.max_peaks(n))   # any k, as in the reference (see get_peak_prominence.py)
    cap = max(5, want)
    rec = np.zeros(1, dtype=record_dtype(cap))
    _cabi.default_context().call("apda_peaks_resolution_f64_host", _p(z.ctypes.data), n, 1, float(fs), _p(0), want, cap,
                                 _p(rec.ctypes.data))
    _cabi.check_record_status(rec)
    return resolution_dicts(rec[0], fs, n)
