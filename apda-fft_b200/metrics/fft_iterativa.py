"""Drop-in for the reference module ``metrics/fft_iterativa.py`` (same names, same call signatures).

Put the directory ``apda-fft_b200/`` ahead of the reference checkout on ``sys.path`` and
``from metrics.fft_iterativa import start_fft`` (reference GT_FFT_v5.py:19) binds this module; the arithmetic
runs in libapda_b200.so on the B200 (fp64, bit-faithful).  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes

import numpy as np

from apda_fft_b200 import _cabi

_p = ctypes.c_void_p


class Spectrum(list):
    """Return type of start_fft: a plain ``list`` ([0, complex, complex, ...]) that also remembers the packed
    complex128 array it was built from, so the pickers can skip re-packing an unmodified spectrum."""

    __slots__ = ("_packed",)

    def __init__(self, values, packed=None):
        super().__init__(values)
        self._packed = packed

    def _drop(self):
        self._packed = None

    def __setitem__(self, key, value):
        self._drop()
        super().__setitem__(key, value)

    def __delitem__(self, key):
        self._drop()
        super().__delitem__(key)

    def __iadd__(self, other):
        self._drop()
        return super().__iadd__(other)

    def __imul__(self, other):
        self._drop()
        return super().__imul__(other)

    def _mutator(name):
        def method(self, *args, **kwargs):
            self._drop()
            return getattr(list, name)(self, *args, **kwargs)
        method.__name__ = name
        return method

    for _name in ("append", "extend", "insert", "pop", "remove", "clear", "sort", "reverse"):
        locals()[_name] = _mutator(_name)
    del _name, _mutator


def pack_spectrum(values) -> np.ndarray:
    """list of complex (bin 0 may be the int 0) -> contiguous complex128 array."""
    packed = getattr(values, "_packed", None)
    if packed is not None and len(packed) == len(values):
        return packed
    return np.ascontiguousarray(np.array(values, dtype=np.complex128))


def remove_dc_component(samples):
    """reference :5-11 - subtract the exact median (device radix select); empty input is returned as is."""
    if not samples:
        return samples
    x = np.ascontiguousarray(samples, dtype=np.float64)
    out = np.empty_like(x)
    _cabi.default_context().call("apda_center_f64_host", _p(x.ctypes.data), x.shape[0], _p(out.ctypes.data))
    return out.tolist()


def pad(lst):
    """reference :13-22 - right-pad with integer zeros to the next power of two (len 0 -> [0])."""
    size = 1
    while size < len(lst):
        size <<= 1
    return lst + [0] * (size - len(lst))


def bit_reversal(x):
    """reference :24-36 - in-place bit-reversal permutation; returns its argument."""
    n = len(x)
    bits = max(n.bit_length() - 1, 0)
    for i in range(n):
        j = int(format(i, f"0{bits}b")[::-1], 2) if bits else 0
        if i < j < n:
            x[i], x[j] = x[j], x[i]
    return x


def fft(x):
    """reference :38-70 - forward unscaled radix-2 DIT FFT of a length-2^k list, in place (returns its argument)."""
    n = len(x)
    if n <= 1:
        return x
    if n & (n - 1):
        raise ValueError("fft: length must be a power of two")
    z = np.ascontiguousarray(np.array(x, dtype=np.complex128))
    out = np.empty_like(z)
    _cabi.default_context().call("apda_fft_c2c_f64_host", _p(z.ctypes.data), 1, n, _p(out.ctypes.data))
    x[:] = out.tolist()
    return x


def start_fft(samples, fs):
    """reference :74-87 - median-centre, zero-pad to 2^k, FFT, bin 0 := 0.  ``fs`` is unused, as in the reference."""
    n_samples = len(samples)
    n = 1
    while n < n_samples:
        n <<= 1
    if n == 1:
        return [0]
    x = np.ascontiguousarray(samples, dtype=np.float64)
    spec = np.empty(n, dtype=np.complex128)
    _cabi.default_context().call("apda_fft_f64_host", _p(x.ctypes.data), n_samples, n_samples, 1, n,
                                 _cabi.CENTER_MEDIAN, _p(spec.ctypes.data))
    values = spec.tolist()
    values[0] = 0
    return Spectrum(values, spec)
