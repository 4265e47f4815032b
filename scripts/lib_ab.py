#!/usr/bin/env python
"""Same-box A/B of library builds (kernel variants): every library in --libs runs in its own process (APDA_LIB) and is
timed on the headline workload's kernels.  Build variants with e.g.
    APDA_NVCC_EXTRA="-DAPDA_K1_MINB=8" APDA_LIB_OUT=$PWD/apda-fft_b200/lib_x.so APDA_OBJ_DIR=$PWD/apda-fft_b200/build_x \
        python apda-fft_b200/build.py

    python scripts/lib_ab.py --libs apda-fft_b200/libapda_b200.so apda-fft_b200/lib_x.so [--windows 400000] [--n 4096]
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
    from apda_fft_b200 import _cabi
    dev = torch.device("cuda:0")
    an = apda_fft_b200.Analyzer(0)
    stream = torch.cuda.current_stream(dev)
    an.use_stream(stream.cuda_stream)
    out = {}
    for dtype, tdt in (("f32", torch.float32), ("f64", torch.float64)):
        if dtype not in args.dtypes:
            continue
        b = args.windows if dtype == "f32" else args.windows // 8
        n = args.n
        x = torch.empty((b, n), dtype=tdt, device=dev)
        spec = torch.empty((b, n, 2), dtype=tdt, device=dev)
        rec = torch.zeros((b, 128), dtype=torch.uint8, device=dev)
        an.synth_device(0, b, n, dtype, x.data_ptr())

        def timed(fn, reps=args.reps):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            best = []
            for _ in range(3):
                a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                for _ in range(reps):
                    fn()
                z.record(stream)
                torch.cuda.synchronize()
                best.append(a.elapsed_time(z) / reps)
            return min(best) * 1e6 / b      # ns per window

        r = {}
        r["k1_median"] = timed(lambda: an.fft_device(x.data_ptr(), b, n, n, dtype, spec.data_ptr()))
        if dtype == "f32":
            r["k1_mean"] = timed(lambda: an.fft_device(x.data_ptr(), b, n, n, dtype, spec.data_ptr(), center=_cabi.CENTER_MEAN))
        an.fft_device(x.data_ptr(), b, n, n, dtype, spec.data_ptr())
        torch.cuda.synchronize()
        r["spec_sha"] = hashlib.sha256(spec[:1000].cpu().numpy().tobytes()).hexdigest()[:12]
        r["k3_flex"] = timed(lambda: an.peaks_device(spec.data_ptr(), b, n, dtype, 125.0, rec.data_ptr(), flexible=True, k=4))
        torch.cuda.synchronize()
        r["rec_flex_sha"] = hashlib.sha256(rec.cpu().numpy().tobytes()).hexdigest()[:12]
        r["k3_rigid"] = timed(lambda: an.peaks_device(spec.data_ptr(), b, n, dtype, 125.0, rec.data_ptr(), flexible=False, k=5))
        torch.cuda.synchronize()
        r["rec_rigid_sha"] = hashlib.sha256(rec.cpu().numpy().tobytes()).hexdigest()[:12]
        if dtype == "f32" and n in (1024, 2048, 4096, 8192):
            r["fused_median"] = timed(lambda: an.analyze_fused_device(x.data_ptr(), b, n, n, 125.0, rec.data_ptr()))
            torch.cuda.synchronize()
            r["rec_fused_sha"] = hashlib.sha256(rec.cpu().numpy().tobytes()).hexdigest()[:12]
            r["fused_mean"] = timed(lambda: an.analyze_fused_device(x.data_ptr(), b, n, n, 125.0, rec.data_ptr(),
                                                                    center=_cabi.CENTER_MEAN))
        out[dtype] = r
        del x, spec, rec
        torch.cuda.empty_cache()
    print("LIBAB " + json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--child", action="store_true")
    ap.add_argument("--libs", nargs="*", default=[])
    ap.add_argument("--windows", type=int, default=400_000)
    ap.add_argument("--n", type=int, default=4096)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--dtypes", default="f32,f64")
    ap.add_argument("--rounds", type=int, default=2, help="every library is run this many times, interleaved")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "lib_ab.json"))
    args = ap.parse_args()
    if args.child:
        return child(args)
    results = {}
    for rnd in range(args.rounds):
        for lib in args.libs:
            env = dict(os.environ, APDA_LIB=os.path.abspath(lib))
            cmd = [sys.executable, os.path.abspath(__file__), "--child", "--windows", str(args.windows), "--n", str(args.n),
                   "--reps", str(args.reps), "--dtypes", args.dtypes]
            res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
            line = [ln for ln in res.stdout.splitlines() if ln.startswith("LIBAB ")]
            results.setdefault(os.path.basename(lib), []).append(json.loads(line[-1][6:]) if line else {"error": res.stderr[-400:]})
    with open(args.out, "w") as fh:
        json.dump(results, fh, indent=1)
    for lib, runs in results.items():
        for dtype in args.dtypes.split(","):
            rows = [r.get(dtype, {}) for r in runs]
            keys = [k for k in rows[0] if not k.endswith("sha")] if rows and rows[0] else []
            print(f"{lib:34s} {dtype} N={args.n} ns/window  " + "  ".join(
                f"{k}={min(r[k] for r in rows if k in r):.2f}" for k in keys)
                + "  " + " ".join(f"{k}={rows[0][k]}" for k in rows[0] if k.endswith("sha")))
        if runs and "error" in runs[0]:
            print(lib, runs[0]["error"])


if __name__ == "__main__":
    main()
