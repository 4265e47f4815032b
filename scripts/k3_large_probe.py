#!/usr/bin/env python
"""Same-box A/B of the large-form picker (K3-large, n >= 2^16): every library in --libs runs in its own process
(APDA_LIB) on the same tone and noise spectra; prints microseconds per window and a hash of the records.

    python scripts/k3_large_probe.py --libs apda-fft_b200/libapda_b200.so apda-fft_b200/lib_old.so
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(args):
    import torch
    import apda_fft_b200
    dev = torch.device("cuda:0")
    an = apda_fft_b200.Analyzer(0)
    stream = torch.cuda.current_stream(dev)
    an.use_stream(stream.cuda_stream)
    out = {}
    for log2n in args.log2n:
        n = 1 << log2n
        for dtype, tdt in (("f32", torch.float32), ("f64", torch.float64)):
            x = torch.empty((1, n), dtype=tdt, device=dev)
            spec = torch.empty((1, n, 2), dtype=tdt, device=dev)
            rec = torch.zeros((1, 128), dtype=torch.uint8, device=dev)
            for kind in ("tones", "noise"):
                if kind == "tones":
                    an.synth_device(0, 1, n, dtype, x.data_ptr())
                else:
                    g = torch.Generator(device=dev).manual_seed(log2n)
                    x.copy_(torch.randn((1, n), generator=g, device=dev, dtype=torch.float32).to(tdt))
                an.fft_device(x.data_ptr(), 1, n, n, dtype, spec.data_ptr())
                for flexible in (True, False):
                    fn = lambda: an.peaks_device(spec.data_ptr(), 1, n, dtype, 250.0, rec.data_ptr(), flexible=flexible)
                    for _ in range(3):
                        fn()
                    torch.cuda.synchronize()
                    best = []
                    for _ in range(3):
                        a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        a.record(stream)
                        for _ in range(args.reps):
                            fn()
                        z.record(stream)
                        torch.cuda.synchronize()
                        best.append(a.elapsed_time(z) / args.reps)
                    r = rec.cpu().numpy().tobytes()
                    key = f"2^{log2n}_{dtype}_{kind}_{'flex' if flexible else 'rigid'}"
                    out[key] = [round(min(best) * 1e3, 1), hashlib.sha256(r).hexdigest()[:10]]
    print("RESULT " + json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--libs", nargs="+")
    ap.add_argument("--log2n", nargs="+", type=int, default=[16, 20, 22, 24])
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--child", action="store_true")
    args = ap.parse_args()
    if args.child:
        return child(args)
    res = {}
    for lib in args.libs:
        env = dict(os.environ, APDA_LIB=os.path.abspath(lib))
        out = subprocess.check_output([sys.executable, __file__, "--child", "--reps", str(args.reps), "--log2n"]
                                      + [str(v) for v in args.log2n], env=env, text=True)
        res[lib] = json.loads([ln for ln in out.splitlines() if ln.startswith("RESULT ")][0][7:])
    keys = list(next(iter(res.values())))
    print(f"{'case':32s}" + "".join(f"{os.path.basename(lib):>34s}" for lib in args.libs))
    for k in keys:
        print(f"{k:32s}" + "".join(f"{res[lib][k][0]:>20.1f} us {res[lib][k][1]:>10s}" for lib in args.libs))


if __name__ == "__main__":
    main()
