"""Batched API over the C ABI: numpy host arrays (H2D/D2H inside the call) or raw device pointers (torch tensors).

    an = Analyzer(device=0)
    recs = an.analyze(samples_f32_or_f64[B, n_samples], fs=125.0, flexible=True)      # numpy structured records
    an.analyze_device(d_samples.data_ptr(), B, n_samples, N, "f32", d_rec.data_ptr()) # everything stays in HBM
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _cabi
from .records import record_dtype

_p = ctypes.c_void_p


def next_pow2(n: int) -> int:
    p = 1
    while p < n:
        p *= 2
    return p


def _suffix(dtype) -> str:
    dt = np.dtype(dtype)
    if dt == np.float64:
        return "f64"
    if dt == np.float32:
        return "f32"
    raise TypeError(f"unsupported dtype {dt}: the kernels compute in float32 or float64")


class Analyzer:
    def __init__(self, device: int = 0, ctx: _cabi.Context | None = None):
        self.ctx = ctx or _cabi.Context(device)

    # ---- host arrays -----------------------------------------------------------------------------------------
    def fft(self, samples: np.ndarray, n_fft: int | None = None, center: int = _cabi.CENTER_MEDIAN) -> np.ndarray:
        """[B, n_samples] real -> [B, N] complex (bin 0 == 0); start_fft of every row."""
        x = np.ascontiguousarray(np.atleast_2d(samples))
        sfx = _suffix(x.dtype)
        b, ns = x.shape
        n = next_pow2(ns) if n_fft is None else int(n_fft)
        out = np.empty((b, n), dtype=np.complex128 if sfx == "f64" else np.complex64)
        self.ctx.call(f"apda_fft_{sfx}_host", _p(x.ctypes.data), ns, ns, b, n, center, _p(out.ctypes.data))
        return out

    def fft_c2c(self, z: np.ndarray) -> np.ndarray:
        z = np.ascontiguousarray(np.atleast_2d(z), dtype=np.complex128)
        out = np.empty_like(z)
        self.ctx.call("apda_fft_c2c_f64_host", _p(z.ctypes.data), z.shape[0], z.shape[1], _p(out.ctypes.data))
        return out

    def peaks(self, spectra: np.ndarray, fs, flexible: bool = True, k: int | None = None,
              rec_cap: int | None = None) -> np.ndarray:
        """[B, n] complex spectra -> records[B]; complex64 input runs the fp32 kernels, anything else the fp64 ones."""
        z = np.atleast_2d(spectra)
        sfx = "f32" if z.dtype == np.complex64 else "f64"
        z = np.ascontiguousarray(z, dtype=np.complex64 if sfx == "f32" else np.complex128)
        b, n = z.shape
        k = (4 if flexible else 5) if k is None else int(k)
        cap = max(5, k) if rec_cap is None else int(rec_cap)
        recs = np.zeros(b, dtype=record_dtype(cap))
        fs_scalar, fs_arr = self._fs(fs, b)
        name = f"apda_peaks_prominence_{sfx}_host" if flexible else f"apda_peaks_resolution_{sfx}_host"
        self.ctx.call(name, _p(z.ctypes.data), n, b, fs_scalar, _p(fs_arr.ctypes.data if fs_arr is not None else 0),
                      k, cap, _p(recs.ctypes.data))
        return recs

    def analyze(self, samples: np.ndarray, fs, flexible: bool = True, k: int | None = None,
                n_fft: int | None = None, center: int = _cabi.CENTER_MEDIAN, rec_cap: int | None = None,
                resolve_ties: bool = True) -> np.ndarray:
        """[B, n_samples] real (float32 or float64) -> records[B]; spectra never leave the device.  float32 windows
        whose record carries APDA_STATUS_FP32_TIE are re-run in float64 (``resolve_ties``)."""
        x = np.ascontiguousarray(np.atleast_2d(samples))
        sfx = _suffix(x.dtype)
        b, ns = x.shape
        n = next_pow2(ns) if n_fft is None else int(n_fft)
        k = (4 if flexible else 5) if k is None else int(k)
        cap = max(5, k) if rec_cap is None else int(rec_cap)
        recs = np.zeros(b, dtype=record_dtype(cap))
        fs_scalar, fs_arr = self._fs(fs, b)
        self.ctx.call(f"apda_analyze_{sfx}_host", _p(x.ctypes.data), ns, ns, b, n, center, int(bool(flexible)),
                      fs_scalar, _p(fs_arr.ctypes.data if fs_arr is not None else 0), k, cap, _p(recs.ctypes.data))
        if sfx == "f32" and resolve_ties:
            self._resolve_fp32_ties(x, recs, fs, flexible, k, n, cap)
        return recs

    def _resolve_fp32_ties(self, x, recs, fs, flexible, k, n, cap) -> int:
        """Windows flagged APDA_STATUS_FP32_TIE (two equal fp32 magnitudes at a peak top: no strict local maximum, so
        the fp32 picker dropped a peak the fp64 reference reports) are re-run through the fp64 pipeline on the same
        samples and their records replaced.  About one window in 10^6 of the fleet generator."""
        tied = np.flatnonzero(recs["status"] & _cabi.STATUS_FP32_TIE)
        if tied.size:
            fs_sel = fs if np.isscalar(fs) else np.asarray(fs, dtype=np.float64)[tied]
            recs[tied] = self.analyze(x[tied].astype(np.float64), fs_sel, flexible=flexible, k=k, n_fft=n, rec_cap=cap)
        return int(tied.size)

    def analyze_fused(self, samples: np.ndarray, fs, flexible: bool = True, k: int | None = None,
                      center: int = _cabi.CENTER_MEDIAN) -> np.ndarray:
        """[B, N] float32, N in {1024, 2048, 4096, 8192} -> records[B] through the fused window->record kernel."""
        x = np.ascontiguousarray(np.atleast_2d(samples), dtype=np.float32)
        b, ns = x.shape
        n = next_pow2(ns)
        k = (4 if flexible else 5) if k is None else int(k)
        recs = np.zeros(b, dtype=record_dtype(5))
        fs_scalar, fs_arr = self._fs(fs, b)
        self.ctx.call("apda_analyze_fused_f32_host", _p(x.ctypes.data), ns, ns, b, n, center, int(bool(flexible)), fs_scalar,
                      _p(fs_arr.ctypes.data if fs_arr is not None else 0), k, 5, _p(recs.ctypes.data))
        return recs

    def analyze_fused_device(self, d_samples: int, batch: int, n_samples: int, n_fft: int, fs: float, d_rec: int,
                             flexible: bool = True, k: int = 4, center: int = _cabi.CENTER_MEDIAN, d_fs: int = 0,
                             ld: int | None = None) -> None:
        self.ctx.call("apda_analyze_fused_f32_dev", _p(d_samples), n_samples, ld or n_samples, batch, n_fft, center,
                      int(bool(flexible)), float(fs), _p(d_fs), k, 5, _p(d_rec))

    # ---- wire-format ingest (16-bit samples, high byte first) ----------------------------------------------------------
    def decode_wire16(self, payload: np.ndarray, first_value) -> tuple[np.ndarray, np.ndarray]:
        """uint8[B, 2*n] payload rows + baseline per window -> (float64[B, n] compacted samples, int32[B] counts)."""
        pay = np.ascontiguousarray(np.atleast_2d(payload), dtype=np.uint8)
        b, nb = pay.shape
        n = nb // 2
        fv = np.ascontiguousarray(np.broadcast_to(np.asarray(first_value, dtype=np.float64), (b,)))
        out = np.zeros((b, n), dtype=np.float64)
        nv = np.zeros(b, dtype=np.int32)
        self.ctx.call("apda_decode_wire16_f64_host", _p(pay.ctypes.data), n, nb, b, _p(fv.ctypes.data), _p(out.ctypes.data),
                      _p(nv.ctypes.data))
        return out, nv

    def analyze_wire16(self, payload: np.ndarray, first_value, fs, n_fft: int | None = None, dtype: str = "f64",
                       flexible: bool = True, k: int | None = None, strict: bool = False) -> np.ndarray:
        """uint8[B, 2*n] payload rows -> records[B] (decode, drop non-finite, centre, FFT, pick; all on the device).
        Windows that lost samples keep their record but carry status bits (APDA_STATUS_OTHER_LENGTH: the reference
        would have transformed them at their own padded length; APDA_STATUS_EMPTY: its pickers raise): inspect
        ``recs["status"]``, or pass ``strict=True`` to raise when any window is not reference-equivalent."""
        pay = np.ascontiguousarray(np.atleast_2d(payload), dtype=np.uint8)
        b, nb = pay.shape
        n = nb // 2
        nfft = next_pow2(n) if n_fft is None else int(n_fft)
        fv = np.ascontiguousarray(np.broadcast_to(np.asarray(first_value, dtype=np.float64), (b,)))
        k = (4 if flexible else 5) if k is None else int(k)
        cap = max(5, k)
        recs = np.zeros(b, dtype=record_dtype(cap))
        fs_scalar, fs_arr = self._fs(fs, b)
        self.ctx.call(f"apda_analyze_wire16_{dtype}_host", _p(pay.ctypes.data), n, nb, b, _p(fv.ctypes.data), nfft,
                      _cabi.CENTER_MEDIAN, int(bool(flexible)), fs_scalar,
                      _p(fs_arr.ctypes.data if fs_arr is not None else 0), k, cap, _p(recs.ctypes.data))
        if strict:
            _cabi.check_record_status(recs)
        return recs

    def analyze_host_ptr(self, h_ptr: int, batch: int, n_samples: int, n_fft: int, dtype: str, fs: float,
                         h_rec_ptr: int, flexible: bool = True, k: int = 4, rec_cap: int = 5,
                         center: int = _cabi.CENTER_MEDIAN) -> None:
        """Same as analyze() on caller-owned (ideally pinned) host buffers given by address."""
        self.ctx.call(f"apda_analyze_{dtype}_host", _p(h_ptr), n_samples, n_samples, batch, n_fft, center,
                      int(bool(flexible)), float(fs), _p(0), k, rec_cap, _p(h_rec_ptr))

    # ---- device pointers (torch tensors: pass .data_ptr()) ---------------------------------------------------------
    def use_stream(self, cuda_stream: int | None) -> None:
        self.ctx.set_stream(cuda_stream)

    def fft_device(self, d_samples: int, batch: int, n_samples: int, n_fft: int, dtype: str, d_spec: int,
                   center: int = _cabi.CENTER_MEDIAN, ld: int | None = None) -> None:
        self.ctx.call(f"apda_fft_{dtype}_dev", _p(d_samples), n_samples, ld or n_samples, batch, n_fft, center, _p(d_spec))

    def peaks_device(self, d_spec: int, batch: int, n: int, dtype: str, fs: float, d_rec: int, flexible: bool = True,
                     k: int = 4, rec_cap: int = 5, d_fs: int = 0) -> None:
        kind = "prominence" if flexible else "resolution"
        self.ctx.call(f"apda_peaks_{kind}_{dtype}_dev", _p(d_spec), n, batch, float(fs), _p(d_fs), k, rec_cap, _p(d_rec))

    def analyze_device(self, d_samples: int, batch: int, n_samples: int, n_fft: int, dtype: str, fs: float, d_rec: int,
                       flexible: bool = True, k: int = 4, rec_cap: int = 5, center: int = _cabi.CENTER_MEDIAN,
                       d_spec_ws: int = 0, d_fs: int = 0, ld: int | None = None) -> None:
        self.ctx.call(f"apda_analyze_{dtype}_dev", _p(d_samples), n_samples, ld or n_samples, batch, n_fft, center,
                      int(bool(flexible)), float(fs), _p(d_fs), k, rec_cap, _p(d_spec_ws), _p(d_rec))

    def synth_device(self, first_window: int, count: int, n: int, dtype: str, d_out: int, seed: int = 42,
                     on_bin: bool = False) -> None:
        self.ctx.call(f"apda_synth_{dtype}_dev", first_window, count, n, seed, int(on_bin), _p(d_out))

    def sync(self) -> None:
        self.ctx.sync()

    def launch_count(self) -> int:
        return self.ctx.launch_count()

    @staticmethod
    def _fs(fs, batch: int):
        if np.isscalar(fs):
            return float(fs), None
        arr = np.ascontiguousarray(fs, dtype=np.float64)
        if arr.shape != (batch,):
            raise ValueError(f"fs must be a scalar or have shape ({batch},)")
        return float(arr[0]) if batch else 0.0, arr


def multi_analyze(contexts, samples: np.ndarray, fs, flexible: bool = True, k: int | None = None,
                  n_fft: int | None = None, center: int = _cabi.CENTER_MEDIAN) -> np.ndarray:
    """One process, several GPUs: [B, n_samples] host windows sharded contiguously over ``contexts`` (``_cabi.Context``
    objects of different devices), every device writing its rows of the returned record table
    (apda_multi_analyze_*_host: one host thread per context, no collective)."""
    x = np.ascontiguousarray(np.atleast_2d(samples))
    sfx = _suffix(x.dtype)
    b, ns = x.shape
    n = next_pow2(ns) if n_fft is None else int(n_fft)
    k = (4 if flexible else 5) if k is None else int(k)
    cap = max(5, k)
    recs = np.zeros(b, dtype=record_dtype(cap))
    fs_scalar, fs_arr = Analyzer._fs(fs, b)
    handles = (ctypes.c_void_p * len(contexts))(*[c.handle for c in contexts])
    fn = getattr(_cabi.load(), f"apda_multi_analyze_{sfx}_host")
    _cabi.check(fn(handles, len(contexts), _p(x.ctypes.data), ns, ns, b, n, center, int(bool(flexible)), fs_scalar,
                   _p(fs_arr.ctypes.data if fs_arr is not None else 0), k, cap, _p(recs.ctypes.data)))
    return recs

