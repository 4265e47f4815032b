#!/usr/bin/env python
"""K1 alone, per centring mode:  python scripts/time_k1.py {f32,f64} N windows"""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "apda-fft_b200"))
import torch
import apda_fft_b200
from apda_fft_b200 import _cabi

dtype, n, windows = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
dev = torch.device("cuda:0")
an = apda_fft_b200.Analyzer(0)
stream = torch.cuda.current_stream(dev)
an.use_stream(stream.cuda_stream)
tdt = torch.float32 if dtype == "f32" else torch.float64
x = torch.empty((windows, n), dtype=tdt, device=dev)
spec = torch.empty((windows, n, 2), dtype=tdt, device=dev)
an.synth_device(0, windows, n, dtype, x.data_ptr())
modes = [("median", _cabi.CENTER_MEDIAN), ("none", _cabi.CENTER_NONE)] + ([("mean", _cabi.CENTER_MEAN)] if dtype == "f32" else [])
for name, c in modes:
    ts = []
    for i in range(8):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        an.fft_device(x.data_ptr(), windows, n, n, dtype, spec.data_ptr(), center=c)
        b.record(stream)
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(a.elapsed_time(b))
    s = 4 if dtype == "f32" else 8
    ms = statistics.median(ts)
    print(f"{dtype} N={n} {windows} windows, centring {name}: {ms:.4f} ms, {3 * s * n * windows / ms / 1e6:.0f} GB/s")
