#!/usr/bin/env python
"""One launch of every kernel family of the round-2 report, for ncu (launch list and --set full captures):

    python scripts/prof_r2.py [what ...]      what in: f32 fused f64 k2   (default: all)

f32   K1 / K3 on 200k windows x 4096 fp32 (flexible and rigid pickers)
fused the fused window->record kernel on the same windows (exact median)
f64   K1 / K3 on 20k windows x 4096 fp64 and 10k x 8192 fp64
k2    one transform of 2^24 samples, fp32 and fp64 (median + passes), and the large-window picker
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import apda_fft_b200  # noqa: E402

what = set(sys.argv[1:]) or {"f32", "fused", "f64", "k2"}
dev = torch.device("cuda:0")
an = apda_fft_b200.Analyzer(0)
an.use_stream(torch.cuda.current_stream(dev).cuda_stream)


def batch(windows, n, dtype, fused=False):
    tdt = torch.float32 if dtype == "f32" else torch.float64
    x = torch.empty((windows, n), dtype=tdt, device=dev)
    spec = torch.empty((windows, n, 2), dtype=tdt, device=dev)
    rec = torch.zeros((windows, 128), dtype=torch.uint8, device=dev)
    an.synth_device(0, windows, n, dtype, x.data_ptr())
    if fused:
        an.analyze_fused_device(x.data_ptr(), windows, n, n, 125.0, rec.data_ptr())
    else:
        an.fft_device(x.data_ptr(), windows, n, n, dtype, spec.data_ptr())
        an.peaks_device(spec.data_ptr(), windows, n, dtype, 125.0, rec.data_ptr(), flexible=True, k=4)
        an.peaks_device(spec.data_ptr(), windows, n, dtype, 125.0, rec.data_ptr(), flexible=False, k=5)
    torch.cuda.synchronize()


if "f32" in what:
    batch(200_000, 4096, "f32")
if "fused" in what:
    batch(200_000, 4096, "f32", fused=True)
if "f64" in what:
    batch(20_000, 4096, "f64")
    batch(10_000, 8192, "f64")
if "k2" in what:
    n = 1 << 24
    for dtype, tdt in (("f32", torch.float32), ("f64", torch.float64)):
        i = torch.arange(n, dtype=torch.float64, device=dev)
        x = (0.5 * torch.sin(2 * torch.pi * 101.6 * i / n) + 0.3 * torch.sin(2 * torch.pi * 252.4 * i / n + 0.3)
             + 0.2 * torch.sin(2 * torch.pi * 498.0 * i / n + 1.1) + 0.125).to(tdt)
        del i
        spec = torch.empty((n, 2), dtype=tdt, device=dev)
        rec = torch.zeros((1, 128), dtype=torch.uint8, device=dev)
        an.fft_device(x.data_ptr(), 1, n, n, dtype, spec.data_ptr())
        an.peaks_device(spec.data_ptr(), 1, n, dtype, 125.0, rec.data_ptr(), flexible=True, k=4)
        torch.cuda.synchronize()
        del x, spec
print("done")
