#!/usr/bin/env python
"""Run one kernel family a few times (for ncu captures):  python scripts/prof_one.py {k1,k3,pipe} {f32,f64} N windows [rigid]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "apda-fft_b200"))
import torch
import apda_fft_b200

what, dtype, n, windows = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
flexible = "rigid" not in sys.argv[5:]
dev = torch.device("cuda:0")
an = apda_fft_b200.Analyzer(0)
an.use_stream(torch.cuda.current_stream(dev).cuda_stream)
tdt = torch.float32 if dtype == "f32" else torch.float64
x = torch.empty((windows, n), dtype=tdt, device=dev)
spec = torch.empty((windows, n, 2), dtype=tdt, device=dev)
rec = torch.zeros((windows, 128), dtype=torch.uint8, device=dev)
an.synth_device(0, windows, n, dtype, x.data_ptr())
for _ in range(3):
    if what in ("k1", "pipe", "k3"):
        an.fft_device(x.data_ptr(), windows, n, n, dtype, spec.data_ptr())
    if what in ("k3", "pipe"):
        an.peaks_device(spec.data_ptr(), windows, n, dtype, 125.0, rec.data_ptr(), flexible=flexible, k=4 if flexible else 5)
torch.cuda.synchronize()
print("done")
