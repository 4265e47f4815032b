#!/usr/bin/env python
"""ns per window of K1 (exact median) on continuous and on ADC-quantised windows (few distinct values, many duplicates of
the median), fp32 and fp64: the counting selection's round count depends on how fast the bracket can close."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import apda_fft_b200
from apda_fft_b200 import _cabi

dev = torch.device("cuda:0")
an = apda_fft_b200.Analyzer(0)
stream = torch.cuda.current_stream(dev)
an.use_stream(stream.cuda_stream)
for n in (1024, 4096, 8192):
    for dtype, tdt, b in (("f32", torch.float32, 200000), ("f64", torch.float64, 40000)):
        x = torch.empty((b, n), dtype=tdt, device=dev)
        spec = torch.empty((b, n, 2), dtype=tdt, device=dev)
        row = []
        for kind in ("tones", "noise", "q50", "q5", "q1", "two-level", "const"):
            g = torch.Generator(device=dev).manual_seed(n)
            if kind == "tones":
                an.synth_device(0, b, n, dtype, x.data_ptr())
            elif kind == "noise":
                x.copy_(torch.randn((b, n), generator=g, device=dev, dtype=torch.float32))
            elif kind.startswith("q"):      # LSB = 1/16384 g, sigma = q LSB, offset 0.98 g (gravity on one axis)
                lv = float(kind[1:])
                r = torch.randn((b, n), generator=g, device=dev, dtype=torch.float32)
                x.copy_((torch.round(r * lv) / 16384.0 + 0.98).to(tdt))
            elif kind == "two-level":
                x.copy_((torch.rand((b, n), generator=g, device=dev) < 0.5).to(tdt) * 0.25)
            else:
                x.fill_(0.5)
            res = {}
            for name, c in (("median", _cabi.CENTER_MEDIAN), ("none", _cabi.CENTER_NONE)):
                fn = lambda: an.fft_device(x.data_ptr(), b, n, n, dtype, spec.data_ptr(), center=c)
                for _ in range(2):
                    fn()
                torch.cuda.synchronize()
                a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                for _ in range(4):
                    fn()
                z.record(stream)
                torch.cuda.synchronize()
                res[name] = a.elapsed_time(z) / 4 * 1e6 / b
            row.append(f"{kind} {res['median']:.2f}")
        print(f"n={n} {dtype} (no centring {res['none']:.2f}): " + "  ".join(row), flush=True)
