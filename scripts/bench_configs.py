#!/usr/bin/env python
"""Secondary measurements: every configuration of BASELINE.json on one B200 (bench.py only times the headline cfg5).

    python scripts/bench_configs.py [--out profiles/configs_rNN.json]

cfg1  single 3-axis window N=1024 fp64 through the drop-in modules (latency, reference-path call sequence)
cfg2  10k windows N=4096 fp64, flexible picker
cfg3  100k windows N=8192 rigid picker, fp32 and fp64 (+ fp32-vs-fp64 tolerance report)
cfg4  single transforms N=2^20, 2^22, 2^24 (multi-pass K2), fp64 and fp32
cfg5  1M windows N=4096 fp32 (the bench.py workload; repeated here for the per-kernel split)
All timings: CUDA events on the launching stream, inputs resident in HBM, 3 warm-ups, median of the repetitions.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "apda-fft_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import apda_fft_b200  # noqa: E402
from apda_fft_b200 import _cabi  # noqa: E402
from apda_fft_b200.records import record_dtype  # noqa: E402

PEAK = 6549.1
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timed(fn, reps=10, warm=3):
    stream = torch.cuda.current_stream()
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    out = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        fn()
        b.record(stream)
        torch.cuda.synchronize()
        out.append(a.elapsed_time(b))
    return statistics.median(out), min(out)


def batch_config(an, dev, windows, n, dtype, flexible, center=_cabi.CENTER_MEDIAN, reps=10):
    tdt = torch.float32 if dtype == "f32" else torch.float64
    s = 4 if dtype == "f32" else 8
    d_x = torch.empty((windows, n), dtype=tdt, device=dev)
    d_spec = torch.empty((windows, n, 2), dtype=tdt, device=dev)
    d_rec = torch.zeros((windows, 128), dtype=torch.uint8, device=dev)
    an.synth_device(0, windows, n, dtype, d_x.data_ptr())
    k = 4 if flexible else 5
    t1, _ = timed(lambda: an.fft_device(d_x.data_ptr(), windows, n, n, dtype, d_spec.data_ptr(), center=center), reps)
    t3, _ = timed(lambda: an.peaks_device(d_spec.data_ptr(), windows, n, dtype, 125.0, d_rec.data_ptr(), flexible=flexible,
                                          k=k), reps)
    tp, _ = timed(lambda: an.analyze_device(d_x.data_ptr(), windows, n, n, dtype, 125.0, d_rec.data_ptr(),
                                            flexible=flexible, k=k, center=center, d_spec_ws=d_spec.data_ptr()), reps)
    recs = d_rec.cpu().numpy().view(record_dtype(5)).reshape(-1)
    b_alg = 4 * s * n + 128
    res = {"windows": windows, "n": n, "dtype": dtype, "picker": "flexible" if flexible else "rigid",
           "centering": "median" if center == _cabi.CENTER_MEDIAN else "mean",
           "k1_ms": t1, "k3_ms": t3, "pipeline_ms": tp, "windows_per_s": windows / (tp * 1e-3),
           "k1_gbs": 3 * s * n * windows / (t1 * 1e-3) / 1e9, "k3_gbs": (s * n + 128) * windows / (t3 * 1e-3) / 1e9,
           "pipeline_gbs": b_alg * windows / (tp * 1e-3) / 1e9,
           "pipeline_frac_of_measured_peak": b_alg * windows / (tp * 1e-3) / 1e9 / PEAK,
           "mean_peaks": float(recs["count"].mean())}
    return res, recs, d_x


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "configs.json"))
    ap.add_argument("--max-log2n", type=int, default=24)
    ap.add_argument("--only", default="", help="comma-separated subset of cfg1,cfg2,cfg3,cfg4,cfg5,ingest")
    args = ap.parse_args()
    only = set(filter(None, args.only.split(",")))

    def want(name):
        return not only or name in only
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    an = apda_fft_b200.Analyzer(0)
    an.use_stream(torch.cuda.current_stream(dev).cuda_stream)
    out = {"peak_gbs": PEAK, "gpu": torch.cuda.get_device_name(0)}

    if want("cfg1"):
        cfg1(out)
    if want("cfg2"):
        out["cfg2"] = batch_config(an, dev, 10_000, 4096, "f64", True)[0]
    if want("cfg3"):
        cfg3(an, dev, out)
    if want("cfg4"):
        cfg4(an, dev, out, args.max_log2n)
    if want("cfg5"):
        # per-kernel split, both centring modes
        for name, center in (("cfg5_median", _cabi.CENTER_MEDIAN), ("cfg5_mean", _cabi.CENTER_MEAN)):
            out[name] = batch_config(an, dev, 1_000_000, 4096, "f32", True, center=center)[0]
            torch.cuda.empty_cache()
        out["cfg5_rigid_median"] = batch_config(an, dev, 1_000_000, 4096, "f32", False)[0]
    if want("ingest"):
        ingest_rows(an, out)

    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out, indent=1))


def cfg1(out):
    # drop-in modules, one 3-axis sensor (three independent single-axis windows, N=1024, fp64)
    from metrics.fft_iterativa import start_fft
    from utils.get_peak_prominence import get_top_peaks_prominence
    axes = [apda_fft_b200.synth.fleet_window(w, 1024).tolist() for w in range(3)]
    for ax in axes:
        get_top_peaks_prominence(start_fft(ax, 125.0), 125.0)
    t0 = time.perf_counter()
    reps = 20
    for _ in range(reps):
        peaks = [get_top_peaks_prominence(start_fft(ax, 125.0), 125.0) for ax in axes]
    out["cfg1"] = {"what": "3 axes x N=1024 fp64 through start_fft + get_top_peaks_prominence (drop-in modules, host lists)",
                   "ms_per_3axis_window": (time.perf_counter() - t0) / reps * 1e3,
                   "idx": [[p["idx"] for p in pk] for pk in peaks]}



def cfg3(an, dev, out):
    # 100k x 8192 rigid, fp32 and fp64 (+ tolerance report)
    r32, rec32, d_x32 = batch_config(an, dev, 100_000, 8192, "f32", False)
    x64 = d_x32.double()
    del d_x32
    d_spec = torch.empty((100_000, 8192, 2), dtype=torch.float64, device=dev)
    d_rec = torch.zeros((100_000, 128), dtype=torch.uint8, device=dev)
    t1, _ = timed(lambda: an.fft_device(x64.data_ptr(), 100_000, 8192, 8192, "f64", d_spec.data_ptr()), 5)
    t3, _ = timed(lambda: an.peaks_device(d_spec.data_ptr(), 100_000, 8192, "f64", 125.0, d_rec.data_ptr(), flexible=False,
                                          k=5), 5)
    rec64 = d_rec.cpu().numpy().view(record_dtype(5)).reshape(-1)
    same_idx = (rec32["count"] == rec64["count"]) & (rec32["pk"]["idx"] == rec64["pk"]["idx"]).all(axis=1)
    live = (rec64["pk"]["idx"] >= 0) & same_idx[:, None]
    rel = np.abs(rec32["pk"]["mag"][live] - rec64["pk"]["mag"][live]) / rec64["pk"]["mag"][live]
    out["cfg3"] = {"f32": r32,
                   "f64": {"k1_ms": t1, "k3_ms": t3, "windows_per_s": 100_000 / ((t1 + t3) * 1e-3),
                           "pipeline_gbs": (4 * 8 * 8192 + 128) * 100_000 / ((t1 + t3) * 1e-3) / 1e9},
                   "f32_vs_f64": {"same inputs": "fp32 samples widened to fp64", "windows": 100_000,
                                  "identical_index_lists": int(same_idx.sum()),
                                  "max_rel_mag_diff_on_identical": float(rel.max()), "tolerance": 1e-5}}
    del x64, d_spec, d_rec
    torch.cuda.empty_cache()



def cfg4(an, dev, out, max_log2n):
    # large single transforms
    cfg4 = []
    for log2n in (20, 22, 24):
        if log2n > max_log2n:
            continue
        n = 1 << log2n
        for dtype, tdt, s in (("f64", torch.float64, 8), ("f32", torch.float32, 4)):
            i = torch.arange(n, dtype=torch.float64, device=dev)
            x = (0.5 * torch.sin(2 * torch.pi * 101.6 * i / n) + 0.3 * torch.sin(2 * torch.pi * 252.4 * i / n + 0.3)
                 + 0.2 * torch.sin(2 * torch.pi * 498.0 * i / n + 1.1)).to(tdt)
            spec = torch.empty((n, 2), dtype=tdt, device=dev)
            rec = torch.zeros((1, 128), dtype=torch.uint8, device=dev)
            an.fft_device(x.data_ptr(), 1, n, n, dtype, spec.data_ptr())          # builds the twiddle table
            torch.cuda.synchronize()
            tf, _ = timed(lambda: an.fft_device(x.data_ptr(), 1, n, n, dtype, spec.data_ptr()), 5, 2)
            tn, _ = timed(lambda: an.fft_device(x.data_ptr(), 1, n, n, dtype, spec.data_ptr(), center=_cabi.CENTER_NONE), 5, 2)
            tk, _ = timed(lambda: an.peaks_device(spec.data_ptr(), 1, n, dtype, 250.0, rec.data_ptr(), flexible=True), 3, 1)
            r = rec.cpu().numpy().view(record_dtype(5)).reshape(-1)[0]
            cfg4.append({"log2n": log2n, "dtype": dtype, "fft_ms_with_median": tf, "fft_ms_no_centering": tn,
                         "transforms_per_s": 1e3 / tf, "b_alg_7sN_gbs": 7 * s * n / (tn * 1e-3) / 1e9,
                         "peaks_ms": tk, "peak_idx": [int(v) for v in r["pk"]["idx"][: int(r["count"])]]})
            del x, spec
            torch.cuda.empty_cache()
    out["cfg4"] = cfg4



def ingest_rows(an, out):
    # ingest rows: text logs and 16-bit wire samples -> records, host buffers in, records out
    import ctypes
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cases
    nlog, ns = 4096, 4096
    one = cases.log_text(1, ns, False).encode()
    region = one[one.index(b"\n", one.index(b"\n", one.index(b"\n", one.index(b"\n") + 1) + 1) + 1) + 1:]
    text = np.frombuffer(region * nlog, dtype=np.uint8)
    off = (np.arange(nlog + 1, dtype=np.int64) * len(region))
    recs = np.zeros(nlog, dtype=record_dtype(5)); nv = np.zeros(nlog, dtype=np.int32); fl = np.zeros(nlog, dtype=np.int32)
    p = ctypes.c_void_p

    def text_step(dtype):
        an.ctx.call(f"apda_analyze_text_{dtype}_host", p(text.ctypes.data), p(off.ctypes.data), nlog, ns, ns, 0, 1, 125.0, p(0),
                    4, 5, p(recs.ctypes.data), p(nv.ctypes.data), p(fl.ctypes.data))
    ingest = {}
    for dtype in ("f32", "f64"):
        text_step(dtype)
        t0 = time.perf_counter()
        for _ in range(3):
            text_step(dtype)
        dt = (time.perf_counter() - t0) / 3
        ingest[f"text_logs_{dtype}"] = {"logs": nlog, "samples_per_log": ns, "text_bytes_per_log": len(region),
                                        "logs_per_s": nlog / dt, "text_gbs": nlog * len(region) / dt / 1e9,
                                        "all_valid": bool((nv == ns).all() and (fl == 0).all())}
    # host baseline: the reference loop (split + float()) on one core
    rows = region.decode().splitlines()
    t0 = time.perf_counter()
    for _ in range(3):
        cnt = 0
        for row in rows:
            for tok in row.strip().split(";"):
                if tok:
                    float(tok); cnt += 1
    ingest["python_float_parse_logs_per_s_per_core"] = 3 / (time.perf_counter() - t0)
    out["ingest"] = ingest


if __name__ == "__main__":
    main()
