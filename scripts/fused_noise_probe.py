#!/usr/bin/env python
"""ns per window of the fused window->record kernel and of the two-kernel pipeline on tone and on noise windows."""
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
b = 200000
for n in (1024, 4096, 8192):
    x = torch.empty((b, n), dtype=torch.float32, device=dev)
    rec = torch.zeros((b, 128), dtype=torch.uint8, device=dev)
    rec2 = torch.zeros((b, 128), dtype=torch.uint8, device=dev)
    for kind in ("tones", "noise"):
        if kind == "tones":
            an.synth_device(0, b, n, "f32", x.data_ptr())
        else:
            x.normal_(generator=torch.Generator(device=dev).manual_seed(n))
        row = []
        for flexible in (True, False):
            fused = lambda: an.analyze_fused_device(x.data_ptr(), b, n, n, 125.0, rec.data_ptr(), flexible=flexible)
            pipe = lambda: an.analyze_device(x.data_ptr(), b, n, n, "f32", 125.0, rec2.data_ptr(), flexible=flexible)
            for name, fn in (("fused", fused), ("pipeline", pipe)):
                for _ in range(2):
                    fn()
                torch.cuda.synchronize()
                a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                for _ in range(5):
                    fn()
                z.record(stream)
                torch.cuda.synchronize()
                row.append(f"{name}/{'flex' if flexible else 'rigid'} {a.elapsed_time(z) / 5 * 1e6 / b:6.2f}")
            same = float((rec == rec2).all(dim=1).float().mean())
            row.append(f"same {same:.4f}")
        print(f"n={n} {kind:6s}: " + "  ".join(row), flush=True)
