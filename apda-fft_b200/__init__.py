"""apda-fft_b200: B200-native (sm_100a) implementation of APDA-FFT's spectral hot path.

Layout
  csrc/ + libapda_b200.so   hand-written CUDA kernels behind the C ABI of include/apda_b200.h
  _cabi.py                  ctypes binding (no CPU fallback)
  metrics/, utils/          drop-in mirrors of the reference's modules (same names and call signatures)
  batch.py                  batched host/device API used by bench.py and the fleet sweep
  fleet.py                  one-process-per-GPU sharding + gather of the peak records to rank 0
  synth.py                  synthetic multi-tone windows (host generator)
"""
from . import records, synth  # noqa: F401
from ._cabi import ApdaError, Context, default_context, load  # noqa: F401
from .batch import Analyzer, multi_analyze  # noqa: F401

__all__ = ["Analyzer", "multi_analyze", "ApdaError", "Context", "default_context", "load", "records", "synth"]
