#!/usr/bin/env python
"""ns per window of the windowed pickers (K3-fast + repair path) and of the whole pipeline on NOISE windows, next to the
tone windows the headline uses: how much the candidate-list overflow / repair path costs on noise-dominated data."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, nargs="+", default=[1024, 4096, 8192])
    ap.add_argument("--windows", type=int, default=100000)
    args = ap.parse_args()
    import torch
    import apda_fft_b200
    dev = torch.device("cuda:0")
    an = apda_fft_b200.Analyzer(0)
    stream = torch.cuda.current_stream(dev)
    an.use_stream(stream.cuda_stream)
    for n in args.n:
        for dtype, tdt in (("f32", torch.float32), ("f64", torch.float64)):
            b = args.windows if dtype == "f32" else args.windows // 4
            x = torch.empty((b, n), dtype=tdt, device=dev)
            spec = torch.empty((b, n, 2), dtype=tdt, device=dev)
            rec = torch.zeros((b, 128), dtype=torch.uint8, device=dev)
            for kind in ("tones", "noise", "tones+noise"):
                if kind == "tones":
                    an.synth_device(0, b, n, dtype, x.data_ptr())
                elif kind == "noise":
                    g = torch.Generator(device=dev).manual_seed(n)
                    x.copy_(torch.randn((b, n), generator=g, device=dev, dtype=torch.float32).to(tdt))
                else:
                    an.synth_device(0, b, n, dtype, x.data_ptr())
                    g = torch.Generator(device=dev).manual_seed(n + 1)
                    x.add_(0.05 * torch.randn((b, n), generator=g, device=dev, dtype=torch.float32).to(tdt))
                an.fft_device(x.data_ptr(), b, n, n, dtype, spec.data_ptr())
                row = []
                for flexible in (True, False):
                    fn = lambda: an.peaks_device(spec.data_ptr(), b, n, dtype, 125.0, rec.data_ptr(), flexible=flexible)
                    for _ in range(2):
                        fn()
                    torch.cuda.synchronize()
                    best = []
                    for _ in range(3):
                        a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        a.record(stream)
                        for _ in range(3):
                            fn()
                        z.record(stream)
                        torch.cuda.synchronize()
                        best.append(a.elapsed_time(z) / 3)
                    counts = rec.cpu().numpy()[:, :4].copy().view("<i4").reshape(-1)
                    row.append(f"{'flex' if flexible else 'rigid'} {min(best) * 1e6 / b:8.2f} ns (mean peaks {counts.mean():.2f})")
                print(f"n={n} {dtype} {kind:12s}: " + "   ".join(row), flush=True)


if __name__ == "__main__":
    main()
