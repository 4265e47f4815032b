#!/usr/bin/env python
"""ns per window of K1 for the three centring modes (median / mean / none), fp32 and fp64, at one window length."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, nargs="+", default=[4096, 8192])
    ap.add_argument("--windows", type=int, default=200000)
    args = ap.parse_args()
    import torch
    import apda_fft_b200
    from apda_fft_b200 import _cabi
    dev = torch.device("cuda:0")
    an = apda_fft_b200.Analyzer(0)
    stream = torch.cuda.current_stream(dev)
    an.use_stream(stream.cuda_stream)
    for n in args.n:
        for dtype, tdt in (("f32", torch.float32), ("f64", torch.float64)):
            b = args.windows if dtype == "f32" else args.windows // 4
            x = torch.empty((b, n), dtype=tdt, device=dev)
            spec = torch.empty((b, n, 2), dtype=tdt, device=dev)
            an.synth_device(0, b, n, dtype, x.data_ptr())
            row = []
            for name, c in (("median", _cabi.CENTER_MEDIAN), ("mean", _cabi.CENTER_MEAN), ("none", _cabi.CENTER_NONE)):
                if dtype == "f64" and name == "mean":
                    continue
                fn = lambda: an.fft_device(x.data_ptr(), b, n, n, dtype, spec.data_ptr(), center=c)
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                best = []
                for _ in range(3):
                    a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(stream)
                    for _ in range(5):
                        fn()
                    z.record(stream)
                    torch.cuda.synchronize()
                    best.append(a.elapsed_time(z) / 5)
                row.append(f"{name} {min(best) * 1e6 / b:7.2f} ns")
            print(f"n={n} {dtype}: " + "   ".join(row), flush=True)


if __name__ == "__main__":
    main()
