#!/usr/bin/env python
"""Same-box A/B of K2 (large transforms) pass structures: every variant runs in its own process (the tuning
environment variables of csrc/fft_large.cu are read once per process).

    python scripts/k2_ab.py [--out gpurun_out/k2_ab.json]            # parent: all variants
    python scripts/k2_ab.py --child                                  # one variant (environment already set)

Per variant and size: ms of the FFT alone (APDA_CENTER_NONE), ms with the exact median, SHA-256 of the spectrum (the
fp64 spectra of every variant must be identical: the dataflow graph does not depend on the pass structure).
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

VARIANTS = {
    "default": {},
    "no PDL": {"APDA_PDL": "0"},
    "PDL median only": {"APDA_PDL": "1"},
    "PDL median+head": {"APDA_PDL": "3"},
    "PDL median+tail": {"APDA_PDL": "5"},
    "PDL head+tail": {"APDA_PDL": "6"},
    "f64 512x1 (round 1)": {"APDA_K2_NT64": "512", "APDA_K2_MINB64": "1"},
    "f32 256x2": {"APDA_K2_MINB32": "2"},
    "f32 512x2": {"APDA_K2_NT32": "512", "APDA_K2_MINB32": "2"},
    "f32 q8": {"APDA_K2_QMAX32": "8"},
    "f32 q11": {"APDA_K2_QMAX32": "11"},
    "f32 q12 128KB": {"APDA_K2_QMAX32": "12", "APDA_K2_TILE32": "16384"},
}


def child():
    import torch
    import apda_fft_b200
    from apda_fft_b200 import _cabi
    dev = torch.device("cuda:0")
    an = apda_fft_b200.Analyzer(0)
    stream = torch.cuda.current_stream(dev)
    an.use_stream(stream.cuda_stream)

    def timed(fn, reps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(reps):
            fn()
        z.record(stream)
        torch.cuda.synchronize()
        return a.elapsed_time(z) / reps

    out = {}
    for log2n in (20, 22, 24):
        n = 1 << log2n
        for dtype, tdt in (("f32", torch.float32), ("f64", torch.float64)):
            i = torch.arange(n, dtype=torch.float64, device=dev)
            x = (0.5 * torch.sin(2 * torch.pi * 101.6 * i / n) + 0.3 * torch.sin(2 * torch.pi * 252.4 * i / n + 0.3)
                 + 0.2 * torch.sin(2 * torch.pi * 498.0 * i / n + 1.1) + 0.125).to(tdt)
            del i
            spec = torch.empty((n, 2), dtype=tdt, device=dev)
            try:
                an.fft_device(x.data_ptr(), 1, n, n, dtype, spec.data_ptr())
                torch.cuda.synchronize()
                sha = hashlib.sha256(spec.cpu().numpy().tobytes()).hexdigest()[:16]
                reps = 200 if log2n < 24 else 100
                t_med = timed(lambda: an.fft_device(x.data_ptr(), 1, n, n, dtype, spec.data_ptr()), reps)
                t_fft = timed(lambda: an.fft_device(x.data_ptr(), 1, n, n, dtype, spec.data_ptr(), center=_cabi.CENTER_NONE), reps)
                out[f"2^{log2n}_{dtype}"] = {"ms_fft": t_fft, "ms_with_median": t_med, "sha": sha}
            except Exception as exc:  # noqa: BLE001
                out[f"2^{log2n}_{dtype}"] = {"error": str(exc)[:200]}
            del x, spec
            torch.cuda.empty_cache()
    print("K2AB " + json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--child", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "k2_ab.json"))
    ap.add_argument("--only", default="")
    ap.add_argument("--libs", nargs="*", default=[], help="library builds to compare (APDA_LIB) instead of the env variants")
    args = ap.parse_args()
    if args.child:
        return child()
    results = {}
    variants = {os.path.basename(p): {"APDA_LIB": os.path.abspath(p)} for p in args.libs} if args.libs else VARIANTS
    for name, env in variants.items():
        if args.only and not any(o in name for o in args.only.split(",")):
            continue
        res = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=dict(os.environ, **env),
                             capture_output=True, text=True, timeout=900)
        line = [ln for ln in res.stdout.splitlines() if ln.startswith("K2AB ")]
        results[name] = json.loads(line[-1][5:]) if line else {"error": res.stderr[-500:]}
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as fh:
        json.dump(results, fh, indent=1)
    keys = sorted({k for r in results.values() for k in r if k.startswith("2^")})
    print("variant".ljust(44) + "".join(k.rjust(17) for k in keys))
    for name, r in results.items():
        cells = []
        for k in keys:
            v = r.get(k, {})
            cells.append((f"{v['ms_fft']*1e3:.0f}/{v['ms_with_median']*1e3:.0f}us" if "ms_fft" in v else "ERR").rjust(17))
        print(name.ljust(44) + "".join(cells))
    base = results.get("default", {})
    for name, r in results.items():
        for k in keys:
            if k.endswith("f64") and "sha" in r.get(k, {}) and "sha" in base.get(k, {}) and r[k]["sha"] != base[k]["sha"]:
                print(f"!! fp64 spectrum of '{name}' differs from default at {k}")


if __name__ == "__main__":
    main()
