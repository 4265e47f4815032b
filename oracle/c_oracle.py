"""ORACLE (test infrastructure, never the product path).

ctypes front end of ``oracle/c/fft_oracle.c`` - the batched, bit-faithful C
restatement of the reference FFT front end - plus picker drivers that feed the
C spectra into the scalar picker port (``oracle/ref_port.py``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
legs may import this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from . import ref_port

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "c", "fft_oracle.c")
_LIB = os.path.join(_HERE, "_build", "libfft_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """gcc -O2 -ffp-contract=off -> oracle/_build/libfft_oracle.so (git-ignored; travels with gpurun)."""
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        os.makedirs(os.path.dirname(_LIB), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC",
                               "-o", _LIB, _SRC, "-lm"])
    return _LIB


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(build())
        p = ctypes.c_void_p
        i64 = ctypes.c_int64
        lib.oracle_twiddle_table.argtypes = [i64, p]
        lib.oracle_start_fft_batch.argtypes = [p, i64, i64, i64, i64, p, p]
        lib.oracle_fft_c2c.argtypes = [p, i64, p, p]
        lib.oracle_half_magnitudes.argtypes = [p, i64, p]
        lib.oracle_median.argtypes = [p, i64]
        lib.oracle_median.restype = ctypes.c_double
        _lib = lib
    return _lib


_tw_cache: dict[int, np.ndarray] = {}


def twiddle_table(n: int) -> np.ndarray:
    """complex128[n-1]: stage with half-span h occupies [h-1, 2h-1)."""
    tw = _tw_cache.get(n)
    if tw is None:
        tw = np.zeros(max(n - 1, 1), dtype=np.complex128)
        if n > 1:
            _load().oracle_twiddle_table(n, tw.ctypes.data)
        if n <= (1 << 16):
            _tw_cache[n] = tw
    return tw


def padded_len(n_samples: int) -> int:
    n = 1
    while n < n_samples:
        n *= 2
    return n


def start_fft_batch(samples: np.ndarray, n_fft: int | None = None) -> np.ndarray:
    """float64[B, n_samples] -> complex128[B, N]; row b == ref start_fft(samples[b]) bit for bit (bin 0 = 0+0j)."""
    x = np.ascontiguousarray(samples, dtype=np.float64)
    if x.ndim == 1:
        x = x[None, :]
    b, ns = x.shape
    n = padded_len(ns) if n_fft is None else n_fft
    out = np.empty((b, n), dtype=np.complex128)
    tw = twiddle_table(n)
    _load().oracle_start_fft_batch(x.ctypes.data, ns, ns, b, n, tw.ctypes.data, out.ctypes.data)
    return out


def fft_c2c(x: np.ndarray) -> np.ndarray:
    """complex128[N] -> complex128[N]; == reference fft(list(x))."""
    x = np.ascontiguousarray(x, dtype=np.complex128)
    out = np.empty_like(x)
    tw = twiddle_table(x.shape[0])
    _load().oracle_fft_c2c(x.ctypes.data, x.shape[0], tw.ctypes.data, out.ctypes.data)
    return out


def half_magnitudes(spec_row: np.ndarray) -> np.ndarray:
    spec_row = np.ascontiguousarray(spec_row, dtype=np.complex128)
    half = spec_row.shape[0] // 2
    mags = np.empty(half, dtype=np.float64)
    _load().oracle_half_magnitudes(spec_row.ctypes.data, half, mags.ctypes.data)
    return mags


def _as_ref_list(spec_row: np.ndarray) -> list:
    out = spec_row.tolist()
    out[0] = 0
    return out


def peaks_prominence(spec_row: np.ndarray, fs: float, k: int = 4):
    """Reference-equivalent flexible picker on one spectrum row."""
    return ref_port.top_peaks_prominence(_as_ref_list(spec_row), fs, k)


def peaks_resolution(spec_row: np.ndarray, fs: float, k: int = 5):
    """Reference-equivalent rigid picker on one spectrum row."""
    return ref_port.top_peaks_resolution(_as_ref_list(spec_row), fs, k)


def analyze_window(samples_row, fs: float, flexible: bool = True, k: int | None = None):
    """start_fft + picker for one window through the C FFT (used by the CPU-baseline legs and smoke())."""
    spec = start_fft_batch(np.asarray(samples_row, dtype=np.float64))[0]
    if flexible:
        return peaks_prominence(spec, fs, 4 if k is None else k)
    return peaks_resolution(spec, fs, 5 if k is None else k)
