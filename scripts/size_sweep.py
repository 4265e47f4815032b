#!/usr/bin/env python
"""Pipeline cost per window over the window lengths (start_fft + flexible picker, device resident): which kernels serve
which length, and where the hand-over between the specialised and the general kernels costs throughput."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import apda_fft_b200

dev = torch.device("cuda:0")
an = apda_fft_b200.Analyzer(0)
stream = torch.cuda.current_stream(dev)
an.use_stream(stream.cuda_stream)
print(f"{'N':>8s} {'dtype':>5s} {'windows':>8s} {'fft ns/win':>11s} {'peaks ns/win':>13s} {'ns/sample':>10s}")
for log2n in range(8, 19):
    n = 1 << log2n
    for dtype, tdt, s in (("f32", torch.float32, 4), ("f64", torch.float64, 8)):
        b = max(4, min(200000, (1 << 28) // (n * s)))
        x = torch.empty((b, n), dtype=tdt, device=dev)
        spec = torch.empty((b, n, 2), dtype=tdt, device=dev)
        rec = torch.zeros((b, 128), dtype=torch.uint8, device=dev)
        an.synth_device(0, b, n, dtype, x.data_ptr())
        res = []
        for fn in (lambda: an.fft_device(x.data_ptr(), b, n, n, dtype, spec.data_ptr()),
                   lambda: an.peaks_device(spec.data_ptr(), b, n, dtype, 125.0, rec.data_ptr(), flexible=True)):
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(3):
                fn()
            z.record(stream)
            torch.cuda.synchronize()
            res.append(a.elapsed_time(z) / 3 * 1e6 / b)
        print(f"{n:8d} {dtype:>5s} {b:8d} {res[0]:11.2f} {res[1]:13.2f} {(res[0] + res[1]) / n:10.4f}", flush=True)
        del x, spec, rec
        torch.cuda.empty_cache()
