#!/usr/bin/env python
"""One large transform (for ncu launch lists):  python scripts/prof_large.py {f32,f64} log2n"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "apda-fft_b200"))
import torch
import apda_fft_b200

dtype, n = sys.argv[1], 1 << int(sys.argv[2])
dev = torch.device("cuda:0")
an = apda_fft_b200.Analyzer(0)
an.use_stream(torch.cuda.current_stream(dev).cuda_stream)
tdt = torch.float32 if dtype == "f32" else torch.float64
i = torch.arange(n, dtype=torch.float64, device=dev)
x = (0.5 * torch.sin(2 * torch.pi * 101.6 * i / n) + 0.3 * torch.sin(2 * torch.pi * 252.4 * i / n + 0.3)).to(tdt)
spec = torch.empty((n, 2), dtype=tdt, device=dev)
for _ in range(3):
    an.fft_device(x.data_ptr(), 1, n, n, dtype, spec.data_ptr())
torch.cuda.synchronize()
