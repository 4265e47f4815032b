#!/usr/bin/env python
"""bench.py - FFT+peak windows/sec at N=4096 on 1..8 B200 (BASELINE.json metric), one JSON line on rank 0.

Workload = BASELINE.json configs[4] ("Fleet sweep: 1M windows N=4096 fp32 sharded across 1/2/4/8 B200"): --windows
(default 1 000 000) windows IN TOTAL, rank r of G owning the contiguous shard shard_bounds(total, G, r), so
the multi-GPU numbers are STRONG scaling of the stated configuration.  A "step" is one pass of the hot path over the
shard that is already resident in HBM: K1 (FFT, samples -> N complex bins), K3 (picker, half spectrum -> 128-byte
record) and - for G > 1 - the records reaching rank 0 (K3 stores them straight into rank 0's table over NVLink; the
NCCL gather is the alternative, --gather nccl).  No collective touches the data path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Extra keys of the line: roofline (dominant kernel K1, timed live with CUDA events), pipeline (whole step vs B_alg),
e2e (pinned host buffers through the C ABI, H2D/D2H inside the clock, on the full per-rank shard), e2e_wire16 (same from
the sensors' 16-bit wire samples), h2d_probe (cudaMemcpyAsync-only ceiling of this box at this N), clocks,
cpu_baseline (the untouched reference from oracle/_ref on the host cores, else the oracle port), configs (N = 1 only:
BASELINE.json configs[1..3] - cfg2, cfg3 fp32/fp64, cfg4 2^20/2^22/2^24 fp32/fp64 - each with ms, bytes, frac on SURVEY
8(d)'s byte contract and its own clock samples), variants (other centring / picker, fused kernel, weak scaling).
Exit code 3: the peer-memory record table differed from the NCCL gather (multi-GPU runs check it every time).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "FFT+peak windows/sec at N=4096"
UNIT = "windows/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--windows", type=int, default=1_000_000, help="windows of the fleet sweep IN TOTAL (sharded over the GPUs)")
    ap.add_argument("--n", type=int, default=4096, help="FFT length / samples per window")
    ap.add_argument("--dtype", choices=["f32", "f64"], default="f32")
    ap.add_argument("--picker", choices=["flexible", "rigid"], default="flexible")
    ap.add_argument("--center", choices=["median", "mean"], default="median")
    ap.add_argument("--e2e-windows", type=int, default=0, help="windows per GPU of the host-buffer (e2e) legs; 0 = the whole shard")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--gather", choices=["peer", "nccl"], default="peer",
                    help="N>1: records stored straight into rank 0's table over NVLink peer memory (default), or "
                         "gathered with NCCL in slices overlapped with compute")
    ap.add_argument("--gather-slices", type=int, default=8,
                    help="N>1, --gather nccl: sub-batches per step whose record gather overlaps the next sub-batch's compute")
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="CPU work budget (per core) of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-variants", action="store_true")
    ap.add_argument("--no-configs", action="store_true")
    ap.add_argument("--no-numa", action="store_true", help="do not bind the process to the GPU's NUMA node")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------------
# CPU legs: the reference's own implementation of the path on the host cores
#   kind "reference": the UNMODIFIED reference (metrics/fft_iterativa.py:74-87 start_fft + utils/get_peak_*.py pickers),
#                     from the build-time copy oracle/_ref (made by __graft_entry__.build() where /root/reference exists;
#                     git-ignored, travels with the snapshot);
#   kind "port":      oracle/ref_port.py, only when that copy is absent.  Measured in the build container on the same
#                     windows: the port is 1.5x FASTER than the reference (11.0 vs 16.5 ms per window at N = 4096), so a
#                     ratio against the port understates the ratio against the reference.
# ---------------------------------------------------------------------------------------------------------------
PORT_SPEED_VS_REFERENCE = 1.5


def cpu_kind() -> str:
    from oracle import ref_copy
    return "reference" if ref_copy.available() else "port"


_cpu_fns = None


def _cpu_functions():
    global _cpu_fns
    if _cpu_fns is None:
        from oracle import ref_copy
        if ref_copy.available():
            ref = ref_copy.RefModules()
            _cpu_fns = (ref.start_fft, ref.get_top_peaks_prominence, ref.get_top_peaks_resolution)
        else:
            from oracle import ref_port
            _cpu_fns = (ref_port.start_fft, ref_port.top_peaks_prominence, ref_port.top_peaks_resolution)
    return _cpu_fns


def _cpu_worker(task):
    first, count, n, flexible = task
    import apda_fft_b200.synth as synth
    start_fft, prominence, resolution = _cpu_functions()
    done = 0
    for w in range(first, first + count):
        x = synth.fleet_window(w, n).tolist()
        spec = start_fft(x, 125.0)
        peaks = prominence(spec, 125.0) if flexible else resolution(spec, 125.0)
        done += 1 if peaks is not None else 0
    return done


def _cpu_gen_worker(task):
    first, count, n, _ = task
    import apda_fft_b200.synth as synth
    for w in range(first, first + count):
        synth.fleet_window(w, n).tolist()
    return count


def cpu_throughput(n: int, flexible: bool, windows_per_core: int, cores: int, pool=None):
    """windows/s of start_fft + picker over `cores` worker processes (multiprocessing.Pool, one fixed slice of the same
    synthetic windows per worker); input generation is timed separately and subtracted (the GPU legs also start with
    resident inputs)."""
    import multiprocessing as mp
    own = pool is None
    if own:
        pool = mp.get_context("spawn").Pool(cores)
    try:
        tasks = [(10_000_000 + c * windows_per_core, windows_per_core, n, flexible) for c in range(cores)]
        pool.map(_cpu_worker, [(0, 1, n, flexible)] * cores)               # warm the workers (imports)
        t0 = time.perf_counter()
        pool.map(_cpu_gen_worker, tasks, chunksize=1)
        t_gen = time.perf_counter() - t0
        t0 = time.perf_counter()
        done = sum(pool.map(_cpu_worker, tasks, chunksize=1))
        t_all = time.perf_counter() - t0
    finally:
        if own:
            pool.close()
            pool.join()
    t = max(t_all - t_gen, 1e-9)
    return done / t, done, t


def cpu_baseline_note(kind: str) -> str:
    if kind == "reference":
        return "unmodified reference (oracle/_ref: metrics/fft_iterativa.py start_fft + utils/get_peak_*.py), fp64"
    return (f"oracle/ref_port.py (oracle/_ref absent on this box); the port runs {PORT_SPEED_VS_REFERENCE}x faster than "
            "the unmodified reference (measured in the build container), fp64")


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path on all host cores, same metric/config; each
    step is a bounded sample of the workload (>= 64 windows per worker, and >= 5 s of wall clock over the run)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0)) or 1
    flexible = args.picker == "flexible"
    kind = cpu_kind()
    ms_per_window = 16.5 if kind == "reference" else 11.0
    per_core = max(64, math.ceil(5.0 / (ms_per_window * 1e-3) / max(args.steps, 1)))
    pool = mp.get_context("spawn").Pool(cores)
    try:
        for _ in range(args.warmup):
            cpu_throughput(args.n, flexible, 2, cores, pool)
        total_t, total_w = 0.0, 0
        for _ in range(args.steps):
            _, done, t = cpu_throughput(args.n, flexible, per_core, cores, pool)
            total_t += t
            total_w += done
    finally:
        pool.close()
        pool.join()
    value = total_w / total_t
    sample = (f"{per_core * cores} windows of the fleet generator per step ({per_core} per worker, {cores} workers), "
              f"{total_t:.1f} s of wall clock in {args.steps} steps; {cpu_baseline_note(kind)}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total_t / max(args.steps, 1), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, max(args.gpus, 1)),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    per = -(-args.windows // world)
    return {
        "workload": f"fleet sweep (BASELINE configs[4]): {args.windows} windows x N={args.n} {args.dtype} in total, sharded "
                    f"over {world} GPU(s) ({per} per GPU), {args.picker} picker (k={'4' if args.picker == 'flexible' else '5'}), "
                    "K1 FFT + K3 peaks + 128 B records to rank 0",
        "windows_total": args.windows, "windows_per_gpu": per, "n_fft": args.n, "picker": args.picker,
        "centering": args.center,
        "parallelism": f"batch-sharded x{world}, " + (
            "records stay on the one GPU" if world == 1 else
            "K3 stores its records into rank 0's table over NVLink peer memory (no collective)" if args.gather == "peer"
            else f"records gathered to rank 0 with NCCL in {args.gather_slices} slices overlapped with compute"),
        "l2": "inputs (windows*N*s bytes per GPU) and spectra far exceed the 126 MB L2; no flush needed",
    }


# ---------------------------------------------------------------------------------------------------------------
# clocks: one nvidia-smi sampler for the whole run, summarised per time window
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        self.index = index
        self.rows = []          # (host time, line)
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def window(self, t0: float, t1: float):
        """Summary of the samples taken in [t0, t1] (host clock); the region should last >= ~0.3 s to hold a few."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvidia-smi unavailable"]}
        deadline = time.time() + 0.3
        while time.time() < deadline and not any(t >= t1 for t, _ in self.rows):
            time.sleep(0.02)            # let the sample that closes the window arrive
        sm, mx, reasons = [], [], set()
        for t, row in list(self.rows):
            if t < t0 or t > t1 + 0.06:
                continue
            parts = [p.strip() for p in row.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(self.NAMES, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()


class NullSampler:
    def window(self, t0, t1):
        return None

    def stop(self):
        pass


# ---------------------------------------------------------------------------------------------------------------
# NUMA: pin this rank (and hence its first-touch pinned buffers) to the node its GPU hangs off
# ---------------------------------------------------------------------------------------------------------------
_ORIG_AFFINITY = None


def bind_to_gpu_numa_node(local: int) -> dict:
    global _ORIG_AFFINITY
    info = {"node": None, "cpus": None}
    try:
        _ORIG_AFFINITY = os.sched_getaffinity(0)
        import torch
        p = torch.cuda.get_device_properties(local)
        bdf = f"{getattr(p, 'pci_domain_id', 0):04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as fh:
            node = int(fh.read().strip())
        info["pci"] = bdf
        if node < 0:
            info["note"] = "the platform reports no NUMA node for the GPU; process left unbound"
            return info
        with open(f"/sys/devices/system/node/node{node}/cpulist") as fh:
            cpus = set()
            for part in fh.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            info.update(node=node, cpus=len(allowed))
        else:
            info["note"] = f"node {node} has no CPU this process may run on"
    except Exception as exc:  # noqa: BLE001 - binding is best effort, the numbers say what happened
        info["note"] = f"not bound ({type(exc).__name__}: {exc})"
    return info


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------
def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel_key: str, windows: int):
    """DRAM bytes (read + write) of one launch of the dominant kernel, from the committed `ncu --set full` capture
    (profiles/traffic.json holds bytes per window of that capture; scaled to this run's windows per launch)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            entry = json.load(fh)[kernel_key]
        return entry["dram_bytes_per_window"] * windows
    except Exception:
        return None


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    numa = {"node": None, "note": "--no-numa"} if args.no_numa else bind_to_gpu_numa_node(local)

    import apda_fft_b200
    from apda_fft_b200 import _cabi
    from apda_fft_b200.fleet import PeerRecordTable, RecordGatherer, gather_records, shard_bounds, shard_capacity
    from apda_fft_b200.records import record_dtype

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    s_bytes = 4 if args.dtype == "f32" else 8
    tdt = torch.float32 if args.dtype == "f32" else torch.float64
    n, total = args.n, args.windows
    lo_w, hi_w = shard_bounds(total, world, rank)
    b = hi_w - lo_w                                   # windows of this rank
    per = shard_capacity(total, world)                # rows every rank owns in the table (the last shard may be shorter)
    flexible = args.picker == "flexible"
    k = 4 if flexible else 5
    center = _cabi.CENTER_MEDIAN if args.center == "median" else _cabi.CENTER_MEAN
    fs = 125.0
    peak, peak_src = measured_peak()

    an = apda_fft_b200.Analyzer(local)
    stream = torch.cuda.current_stream(dev)
    an.use_stream(stream.cuda_stream)
    sampler = ClockSampler(local).start() if rank == 0 else NullSampler()

    def fence():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(values):
        t = torch.tensor(values, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t]

    class Fleet:
        """One sharded workload: `rows` windows on this rank (global ids from `first`), `cap` table rows per rank."""

        def __init__(self, first, rows, cap, want_peer):
            self.rows, self.cap = rows, cap
            self.d_x = torch.empty((max(rows, 1), n), dtype=tdt, device=dev)
            self.d_spec = torch.empty((max(rows, 1), n, 2), dtype=tdt, device=dev)
            self.d_rec = torch.zeros((cap, 128), dtype=torch.uint8, device=dev)
            if rows:
                an.synth_device(first, rows, n, args.dtype, self.d_x.data_ptr())
            torch.cuda.synchronize()
            self.gatherer = RecordGatherer(cap, 128, dev)
            self.peer, self.note, self.step_no = None, None, 0
            if want_peer:
                try:
                    self.peer = PeerRecordTable(an.ctx, cap, 128, dev)
                    ok = torch.ones(1, device=dev)
                except Exception as exc:  # CUDA IPC unavailable / no peer access: every rank falls back together
                    self.note = f"peer table unavailable ({type(exc).__name__}: {exc}); NCCL gather used"
                    ok = torch.zeros(1, device=dev)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)
                if float(ok[0]) < 1.0:
                    if self.peer is not None:
                        self.peer.close()
                    self.peer = None
                    self.note = self.note or "peer table unavailable on another rank; NCCL gather used"
            self.slices = self.gatherer.slices(1 if (self.peer or world == 1) else args.gather_slices)

        def step(self, events=None, center=center, flexible=flexible, use_peer=True):
            """K1 + K3 (+ records to rank 0) of this rank's shard; returns the full table on rank 0."""
            kk = 4 if flexible else 5
            if self.peer is not None and use_peer:
                self.step_no += 1
                ptr = self.peer.begin(self.step_no)
                if self.rows:
                    if events is not None:
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record(stream)
                    an.fft_device(self.d_x.data_ptr(), self.rows, n, n, args.dtype, self.d_spec.data_ptr(), center=center)
                    if events is not None:
                        e1.record(stream)
                        events.append((e0, e1))
                    an.peaks_device(self.d_spec.data_ptr(), self.rows, n, args.dtype, fs, ptr, flexible=flexible, k=kk, rec_cap=5)
                self.peer.signal(self.step_no)
                if self.peer.owner:
                    table = self.peer.wait(self.step_no)
                    self.peer.release(self.step_no)     # nothing consumes the table inside the loop: hand it back at once
                    return table
                return None
            for lo, hi in self.slices:
                hi = min(hi, self.rows)
                if hi > lo:
                    if events is not None:
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record(stream)
                    an.fft_device(self.d_x[lo:].data_ptr(), hi - lo, n, n, args.dtype, self.d_spec[lo:].data_ptr(), center=center)
                    if events is not None:
                        e1.record(stream)
                        events.append((e0, e1))
                    an.peaks_device(self.d_spec[lo:].data_ptr(), hi - lo, n, args.dtype, fs, self.d_rec[lo:].data_ptr(),
                                    flexible=flexible, k=kk, rec_cap=5)
            for lo, hi in self.slices:
                self.gatherer.start(self.d_rec, lo, hi)
            return self.gatherer.finish()

        def timed(self, steps, warm, **kw):
            """(ms per step, K1 ms per step, clocks) - device events, barrier + synchronize on both sides, max over ranks."""
            for _ in range(warm):
                self.step(**kw)
            fence()
            ev = []
            t0 = time.time()
            a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            table = None
            for _ in range(steps):
                table = self.step(events=ev, **kw)
            z.record(stream)
            fence()
            t1 = time.time()
            ms, k1 = reduce_max([a.elapsed_time(z) / steps, sum(p.elapsed_time(q) for p, q in ev) / max(steps, 1)])
            return ms, k1, sampler.window(t0, t1), table

        def close(self):
            if self.peer is not None:
                fence()
                self.peer.close()
                self.peer = None

    # ---- headline: the stated configuration, strong-scaled ------------------------------------------------------------
    want_peer = world > 1 and args.gather == "peer"
    fleet = Fleet(lo_w, b, per, want_peer)
    if want_peer and fleet.peer is None:
        args.gather = "nccl"
    # peer path: a few extra untimed steps so that the NVLink links carrying the (small) record traffic are out of their
    # idle power state before the clock starts
    warm = max(args.warmup, 3) + (5 if fleet.peer is not None else 0)
    launches0 = an.launch_count()
    step_ms, k1_ms, clocks, table = fleet.timed(args.steps, warm)
    launches = (an.launch_count() - launches0) * args.steps // (args.steps + warm)

    # ---- the table of one more step, and its NCCL twin (hard check on every multi-GPU run) -------------------------
    nccl_equal = None
    if world > 1:
        fence()
        table = fleet.step()
        fence()
        mine = table[:total].clone() if rank == 0 else None
        if fleet.peer is not None:
            twin = fleet.step(use_peer=False)
            fence()
            if rank == 0:
                nccl_equal = bool(torch.equal(mine, twin[:total]))
        table = mine
    summary = None
    if rank == 0:
        recs = table[:total].cpu().numpy().view(record_dtype(5)).reshape(-1)
        summary = {"windows_in_table": int(recs.shape[0]), "mean_peaks_per_window": float(recs["count"].mean()),
                   "status_nonzero": int((recs["status"] != 0).sum()),
                   "fp32_tie_windows": int((recs["status"] & _cabi.STATUS_FP32_TIE != 0).sum())}
        if fleet.note:
            summary["note"] = fleet.note
        if fleet.peer is not None:
            summary["peer_table_equals_nccl_gather"] = nccl_equal
            summary["peer_wait_timed_out"] = fleet.peer.timed_out()

    # ---- variants of the same workload (never the headline) ----------------------------------------------------------
    variants = {}

    def variant_line(ms, k1, clk, rows_total):
        return {"value": rows_total / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "k1_ms_per_step": k1, "clocks": clk,
                "pipeline_frac_of_peak": (4 * s_bytes * n + 128) * (rows_total / world) / (ms * 1e-3) / 1e9 / peak}

    if not args.no_variants and world == 1 and args.dtype == "f32":
        other_center = _cabi.CENTER_MEAN if center == _cabi.CENTER_MEDIAN else _cabi.CENTER_MEDIAN
        ms, k1, clk, _ = fleet.timed(args.steps, 3, center=other_center)
        variants["centering_" + ("mean" if other_center == _cabi.CENTER_MEAN else "median")] = dict(
            variant_line(ms, k1, clk, total),
            note="APDA_CENTER_MEAN is the documented opt-in, legal only when n_samples == N (bins >= 1 do not depend on "
                 "the centring constant; bin 0 is zeroed)")
        ms, k1, clk, _ = fleet.timed(args.steps, 3, flexible=not flexible)
        variants["picker_" + ("rigid" if flexible else "flexible")] = variant_line(ms, k1, clk, total)

        # leaner variant (SURVEY 8d "never mix"): fused window->record kernel, its own byte accounting B_min = s*N + 128
        def fused_variant(v_center):
            def run():
                an.analyze_fused_device(fleet.d_x.data_ptr(), b, n, n, fs, fleet.d_rec.data_ptr(), flexible=flexible, k=k,
                                        center=v_center)
            for _ in range(3):
                run()
            fence()
            t0 = time.time()
            a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(args.steps):
                run()
            z.record(stream)
            fence()
            ms = a.elapsed_time(z) / args.steps
            b_min = s_bytes * n + 128
            return {"value": b / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "bytes_per_window": b_min,
                    "achieved_gbs": b_min * b / (ms * 1e-3) / 1e9, "clocks": sampler.window(t0, time.time()),
                    "note": "spectrum never written to HBM; compute-bound, so the HBM fraction is not its yardstick"}
        # leaner variant: the batched pipeline entry point with the library's own spectrum workspace - K1
        # then writes only the bins [0, N/2) the picker reads; byte accounting B_half = 3*s*N + 128
        def half_variant():
            def run():
                an.analyze_device(fleet.d_x.data_ptr(), b, n, n, args.dtype, fs, fleet.d_rec.data_ptr(), flexible=flexible, k=k,
                                  center=center)
            for _ in range(3):
                run()
            fence()
            t0 = time.time()
            a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(args.steps):
                run()
            z.record(stream)
            fence()
            ms = a.elapsed_time(z) / args.steps
            b_half = 3 * s_bytes * n + 128
            return {"value": b / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "bytes_per_window": b_half,
                    "achieved_gbs": b_half * b / (ms * 1e-3) / 1e9, "frac_of_peak": b_half * b / (ms * 1e-3) / 1e9 / peak,
                    "clocks": sampler.window(t0, time.time()),
                    "api": f"apda_analyze_{args.dtype}_dev with d_spec_ws = NULL (exact median, same records as the headline)"}
        if n in (1024, 2048, 4096, 8192):
            variants["half_spectrum_pipeline"] = half_variant()
            variants["fused_kernel_median"] = fused_variant(_cabi.CENTER_MEDIAN)
            variants["fused_kernel_mean"] = fused_variant(_cabi.CENTER_MEAN)
        # the same pipeline on windows of pure noise (a fleet at rest): dozens of hot local maxima per window instead of
        # three tones - the picker's worst case; the windows are regenerated afterwards
        g = torch.Generator(device=dev).manual_seed(1234)
        for lo in range(0, b, 65536):
            fleet.d_x[lo:lo + 65536].normal_(generator=g)
        ms, k1, clk, tab = fleet.timed(args.steps, 3)
        nrec = tab[:total].cpu().numpy().view(record_dtype(5)).reshape(-1)
        variants["noise_windows"] = dict(variant_line(ms, k1, clk, total), mean_peaks_per_window=float(nrec["count"].mean()),
                                         status_nonzero=int((nrec["status"] != 0).sum()),
                                         note="standard-normal samples instead of the three-tone windows; same kernels")
        # ... and on windows as a 16-bit sensor delivers them: the three tones at 20 mg over 13 LSB of noise and a 0.98 g
        # offset, rounded to the ADC's 1/16384 g - a few hundred distinct values per window, the median value repeated
        an.synth_device(lo_w, b, n, args.dtype, fleet.d_x.data_ptr())
        for lo in range(0, b, 65536):
            blk = fleet.d_x[lo:lo + 65536]
            blk.mul_(0.02 * 16384.0).add_(torch.randn(blk.shape, generator=g, device=dev, dtype=blk.dtype), alpha=13.0)
            blk.round_().div_(16384.0).add_(0.98)
        ms, k1, clk, tab = fleet.timed(args.steps, 3)
        qrec = tab[:total].cpu().numpy().view(record_dtype(5)).reshape(-1)
        variants["quantised_windows"] = dict(variant_line(ms, k1, clk, total), mean_peaks_per_window=float(qrec["count"].mean()),
                                             status_nonzero=int((qrec["status"] != 0).sum()),
                                             note="tones + noise rounded to 1/16384 (16-bit sensor words); same kernels")
        an.synth_device(lo_w, b, n, args.dtype, fleet.d_x.data_ptr())
        fleet.step()            # leave the headline configuration's records in d_rec for the e2e comparison
        fence()
    if not args.no_variants and world > 1:
        # weak scaling (round 1's headline): the same `--windows` on EVERY GPU
        weak = Fleet(rank * total, total, total, fleet.peer is not None)
        ms, k1, clk, _ = weak.timed(args.steps, 8)
        variants["weak"] = dict(variant_line(ms, k1, clk, total * world), scaling="weak",
                                note=f"{total} windows per GPU ({total * world} in total): per-GPU work fixed as N grows")
        weak.close()
        del weak
        torch.cuda.empty_cache()

    # ---- e2e: pinned host buffers through the C ABI, copies inside the clock, on the whole shard of every rank ----------
    e2e = e2e_wire = h2d_probe = None
    if not args.no_e2e and b > 0:
        eb = b if args.e2e_windows <= 0 else min(args.e2e_windows, b)
        h_x = torch.empty((eb, n), dtype=tdt).pin_memory()
        h_x.copy_(fleet.d_x[:eb])
        h_rec = torch.zeros((eb, 128), dtype=torch.uint8).pin_memory()
        torch.cuda.synchronize()

        def e2e_step():
            an.analyze_host_ptr(h_x.data_ptr(), eb, n, n, args.dtype, fs, h_rec.data_ptr(), flexible=flexible, k=k,
                                rec_cap=5, center=center)

        def wall(fn, reps):
            fn()
            fence()
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            torch.cuda.synchronize()
            return reduce_max([time.perf_counter() - t0])[0]

        t_e2e = wall(e2e_step, args.e2e_steps)
        if fleet.peer is None:
            local_recs = fleet.d_rec[:eb]
        else:
            an.peaks_device(fleet.d_spec.data_ptr(), eb, n, args.dtype, fs, fleet.d_rec.data_ptr(), flexible=flexible, k=k, rec_cap=5)
            torch.cuda.synchronize()
            local_recs = fleet.d_rec[:eb]
        same = bool((h_rec.numpy() == local_recs.cpu().numpy()).all())
        e2e = {"value": eb * world * args.e2e_steps / t_e2e, "unit": UNIT,
               "h2d_bytes_per_step": eb * n * s_bytes * world, "d2h_bytes_per_step": eb * 128 * world,
               "windows_per_gpu_per_step": eb, "records_equal_device_path": same,
               "h2d_gbs_per_gpu": eb * n * s_bytes * args.e2e_steps / t_e2e / 1e9,
               "api": f"apda_analyze_{args.dtype}_host (pinned host buffers, chunked 2-stream H2D/compute/D2H)"}

        # the ceiling of this box at this N: the same pinned buffer, cudaMemcpyAsync only (one copy per 96 MB chunk, all
        # ranks at once) - what e2e could reach if the kernels were free
        chunk_rows = max(1, (96 << 20) // (n * s_bytes))

        def probe():
            for lo in range(0, eb, chunk_rows):
                hi = min(eb, lo + chunk_rows)
                fleet.d_x[lo:hi].copy_(h_x[lo:hi], non_blocking=True)
        t_probe = wall(probe, args.e2e_steps)
        h2d_probe = {"gbs_per_gpu": eb * n * s_bytes * args.e2e_steps / t_probe / 1e9,
                     "gbs_all_gpus": eb * n * s_bytes * world * args.e2e_steps / t_probe / 1e9,
                     "what": "cudaMemcpyAsync host->device only, same pinned buffer and chunking, all ranks concurrently",
                     "e2e_fraction_of_probe": t_probe / t_e2e}
        del h_x

        # the fleet's documented ingest: the sensors' 16-bit wire samples (2 bytes per sample over PCIe)
        if args.dtype == "f32":
            gen = torch.Generator(device=dev)
            gen.manual_seed(1234 + rank)
            h_pay = torch.empty((eb, 2 * n), dtype=torch.uint8).pin_memory()
            for lo in range(0, eb, 65536):           # finite 16-bit words (exponent 31 cleared), high byte first
                hi = min(eb, lo + 65536)
                wds = torch.randint(0, 1 << 16, (hi - lo, n), device=dev, generator=gen, dtype=torch.int32) & 0xBFFF
                pay = torch.stack(((wds >> 8).to(torch.uint8), (wds & 0xFF).to(torch.uint8)), dim=2).reshape(hi - lo, 2 * n)
                h_pay[lo:hi].copy_(pay)
            h_fv = (torch.rand(eb, dtype=torch.float64) * 2 - 1).pin_memory()
            h_rec2 = torch.zeros((eb, 128), dtype=torch.uint8).pin_memory()
            torch.cuda.synchronize()
            import ctypes

            def wire_step():
                an.ctx.call("apda_analyze_wire16_f32_host", ctypes.c_void_p(h_pay.data_ptr()), n, 2 * n, eb,
                            ctypes.c_void_p(h_fv.data_ptr()), n, _cabi.CENTER_MEDIAN, int(flexible), fs, ctypes.c_void_p(0),
                            k, 5, ctypes.c_void_p(h_rec2.data_ptr()))
            t_w = wall(wire_step, args.e2e_steps)
            e2e_wire = {"value": eb * world * args.e2e_steps / t_w, "unit": UNIT,
                        "h2d_bytes_per_step": eb * (2 * n + 8) * world, "d2h_bytes_per_step": eb * 128 * world,
                        "windows_per_gpu_per_step": eb,
                        "api": "apda_analyze_wire16_f32_host (raw 16-bit sensor samples + baseline; decode, centre, FFT, pick on device)",
                        "data": "random finite 16-bit words (throughput only; parity is covered by tests/golden wire cases)"}
            del h_pay

    # ---- BASELINE.json configs[1..3] on this GPU (N = 1 only; parity of each is tests/test_gpu_full_size.py) -----------
    configs = None
    if world == 1 and not args.no_configs:
        fleet.close()
        del fleet.d_x, fleet.d_spec
        torch.cuda.empty_cache()
        configs = run_configs(an, dev, stream, sampler, peak)

    if rank == 0:
        k1_bytes = 3 * s_bytes * n * b
        b_alg = (4 * s_bytes * n + 128) * b
        value = total / (step_ms * 1e-3)
        k1_gbs = k1_bytes / (k1_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic", "config": workload_config(args, world),
            "roofline": {"bound": "hbm", "kernel": "K1 fft (samples -> N complex bins)", "achieved": k1_gbs,
                         "peak": peak, "unit": "GB/s", "frac": k1_gbs / peak, "peak_source": peak_src,
                         "bytes_per_launch": k1_bytes, "ms_per_launch": k1_ms / len(fleet.slices),
                         "launches_per_step": len(fleet.slices), "windows_per_launch": b // len(fleet.slices),
                         "traffic": ncu_traffic(f"k1_{args.dtype}_n{n}", b // len(fleet.slices)),
                         "traffic_source": "profiles/traffic.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per window x windows per launch)"},
            "pipeline": {"b_alg_bytes_per_window": 4 * s_bytes * n + 128,
                         "achieved_gbs_per_gpu": b_alg / (step_ms * 1e-3) / 1e9,
                         "frac_of_peak": b_alg / (step_ms * 1e-3) / 1e9 / peak,
                         "k1_share_of_step": k1_ms / step_ms},
            "e2e": e2e, "gpu_launches": int(launches) * world, "clocks": clocks, "result_check": summary,
            "variants": variants, "e2e_wire16": e2e_wire, "h2d_probe": h2d_probe, "numa": numa, "configs": configs,
        }
        if world == 1 and not args.no_cpu_baseline:
            if _ORIG_AFFINITY:          # the CPU leg uses every host core again, not only the GPU's NUMA node
                os.sched_setaffinity(0, _ORIG_AFFINITY)
            cores = len(os.sched_getaffinity(0)) or 1
            kind = cpu_kind()
            per_core = max(64, int(args.cpu_seconds / (0.0165 if kind == "reference" else 0.011)))
            rate, done, t = cpu_throughput(args.n, flexible, per_core, cores)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": kind,
                                    "sample": f"{done} windows of the same generator ({per_core} per worker, "
                                              f"multiprocessing.Pool({cores})) in {t:.1f} s; {cpu_baseline_note(kind)}"}
        print(json.dumps(line), flush=True)
    sampler.stop()
    failed = world > 1 and fleet.peer is not None and rank == 0 and nccl_equal is False
    fleet.close()
    if world > 1:
        flag = torch.tensor([1.0 if failed else 0.0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        failed = float(flag[0]) > 0
        dist.barrier()
        dist.destroy_process_group()
    if failed:
        sys.stderr.write("bench.py: the peer-memory record table differs from the NCCL gather\n")
        sys.exit(3)


def run_configs(an, dev, stream, sampler, peak):
    """cfg2, cfg3 (fp32 + fp64) and cfg4 (2^20 / 2^22 / 2^24, fp32 + fp64) of BASELINE.json on one GPU.  Each entry: the
    device time of one pass over the resident input (CUDA events around enough repetitions to last >= ~0.5 s, so the
    clock sampler sees the load), algorithmic bytes on SURVEY 8(d)'s contract, and the fraction of the measured HBM peak."""
    import torch
    from apda_fft_b200 import _cabi
    from apda_fft_b200.records import record_dtype

    def timed(fn, min_s=0.5, max_reps=4000):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        fn()
        z.record(stream)
        torch.cuda.synchronize()
        reps = int(min(max_reps, max(5, math.ceil(min_s * 1e3 / max(a.elapsed_time(z), 1e-3)))))
        t0 = time.time()
        a.record(stream)
        for _ in range(reps):
            fn()
        z.record(stream)
        torch.cuda.synchronize()
        return a.elapsed_time(z) / reps, reps, sampler.window(t0, time.time())

    def batch(windows, n, dtype, flexible):
        tdt = torch.float32 if dtype == "f32" else torch.float64
        s = 4 if dtype == "f32" else 8
        k = 4 if flexible else 5
        d_x = torch.empty((windows, n), dtype=tdt, device=dev)
        d_spec = torch.empty((windows, n, 2), dtype=tdt, device=dev)
        d_rec = torch.zeros((windows, 128), dtype=torch.uint8, device=dev)
        an.synth_device(0, windows, n, dtype, d_x.data_ptr(), on_bin=not flexible)
        k1, _, _ = timed(lambda: an.fft_device(d_x.data_ptr(), windows, n, n, dtype, d_spec.data_ptr()))
        k3, _, _ = timed(lambda: an.peaks_device(d_spec.data_ptr(), windows, n, dtype, 125.0, d_rec.data_ptr(),
                                                 flexible=flexible, k=k))
        ms, reps, clk = timed(lambda: an.analyze_device(d_x.data_ptr(), windows, n, n, dtype, 125.0, d_rec.data_ptr(),
                                                        flexible=flexible, k=k, d_spec_ws=d_spec.data_ptr()))
        recs = d_rec.cpu().numpy().view(record_dtype(5)).reshape(-1)
        bytes_ = (4 * s * n + 128) * windows
        out = {"workload": f"{windows} windows x N={n} {dtype}, {'flexible' if flexible else 'rigid'} picker, exact median",
               "ms": ms, "reps": reps, "k1_ms": k1, "k3_ms": k3, "windows_per_s": windows / (ms * 1e-3),
               "bytes": bytes_, "bytes_contract": "B_alg = 4*s*N + 128 per window (K1 3sN + K3 sN+128)",
               "achieved_gbs": bytes_ / (ms * 1e-3) / 1e9, "frac": bytes_ / (ms * 1e-3) / 1e9 / peak,
               "k1_frac": 3 * s * n * windows / (k1 * 1e-3) / 1e9 / peak,
               "k3_frac": (s * n + 128) * windows / (k3 * 1e-3) / 1e9 / peak,
               "mean_peaks_per_window": float(recs["count"].mean()), "status_nonzero": int((recs["status"] != 0).sum()),
               "clocks": clk}
        del d_x, d_spec, d_rec
        torch.cuda.empty_cache()
        return out, recs

    out = {"peak_gbs": peak, "note": "inputs resident in HBM; L2 (126 MB) is far smaller than every working set except "
                                     "cfg4 2^20 (fp32: 12 MB, fp64: 25 MB), which is marked l2_resident"}
    out["cfg1_3axis_n1024_f64_dropin"] = cfg1_latency()
    out["cfg2_10k_n4096_f64_flexible"], _ = batch(10_000, 4096, "f64", True)
    out["cfg3_100k_n8192_f32_rigid"], r32 = batch(100_000, 8192, "f32", False)
    out["cfg3_100k_n8192_f64_rigid"], r64 = batch(100_000, 8192, "f64", False)
    import numpy as np
    same = (r32["count"] == r64["count"]) & (r32["pk"]["idx"] == r64["pk"]["idx"]).all(axis=1)
    live = (r64["pk"]["idx"] >= 0) & same[:, None]
    rel = np.abs(r32["pk"]["mag"][live] - r64["pk"]["mag"][live]) / r64["pk"]["mag"][live]
    out["cfg3_f32_vs_f64"] = {"windows": 100_000, "identical_index_lists": int(same.sum()),
                              "max_rel_mag_diff": float(rel.max()) if rel.size else None, "tolerance": 1e-5,
                              "note": "same generator (on-bin tones), fp32 and fp64 instances"}
    for log2n in (20, 22, 24):
        nn = 1 << log2n
        for dtype, tdt, s in (("f32", torch.float32, 4), ("f64", torch.float64, 8)):
            i = torch.arange(nn, dtype=torch.float64, device=dev)
            x = (0.5 * torch.sin(2 * torch.pi * 101.6 * i / nn) + 0.3 * torch.sin(2 * torch.pi * 252.4 * i / nn + 0.3)
                 + 0.2 * torch.sin(2 * torch.pi * 498.0 * i / nn + 1.1)).to(tdt)
            del i
            spec = torch.empty((nn, 2), dtype=tdt, device=dev)
            rec = torch.zeros((1, 128), dtype=torch.uint8, device=dev)
            an.fft_device(x.data_ptr(), 1, nn, nn, dtype, spec.data_ptr())          # builds the twiddle table
            torch.cuda.synchronize()
            ms, reps, clk = timed(lambda: an.fft_device(x.data_ptr(), 1, nn, nn, dtype, spec.data_ptr()))
            ms_fft, _, _ = timed(lambda: an.fft_device(x.data_ptr(), 1, nn, nn, dtype, spec.data_ptr(),
                                                       center=_cabi.CENTER_NONE))
            ms_pk, _, _ = timed(lambda: an.peaks_device(spec.data_ptr(), 1, nn, dtype, 250.0, rec.data_ptr(), flexible=True),
                                min_s=0.2)
            r = rec.cpu().numpy().view(record_dtype(5)).reshape(-1)[0]
            bytes_ = 7 * s * nn
            out[f"cfg4_2^{log2n}_{dtype}"] = {
                "workload": f"one transform N=2^{log2n} {dtype} (K2 multi-pass), exact median + FFT",
                "ms": ms, "reps": reps, "ms_fft_no_centering": ms_fft, "ms_peaks": ms_pk, "transforms_per_s": 1e3 / ms,
                "bytes": bytes_, "bytes_contract": "7*s*N per transform (2-pass FFT: read sN + write 2sN, read 2sN + write 2sN)",
                "achieved_gbs": bytes_ / (ms * 1e-3) / 1e9, "frac": bytes_ / (ms * 1e-3) / 1e9 / peak,
                "frac_fft_no_centering": bytes_ / (ms_fft * 1e-3) / 1e9 / peak,
                "l2_resident": 2 * s * nn <= (126 << 20),
                "peak_idx": [int(v) for v in r["pk"]["idx"][: int(r["count"])]], "clocks": clk}
            del x, spec
            torch.cuda.empty_cache()
    return out


def cfg1_latency():
    """BASELINE configs[0]: one 3-axis sensor (three independent single-axis windows, N = 1024, fp64) through the drop-in
    modules exactly as the gateway calls them (start_fft then get_top_peaks_prominence on host lists), and the unmodified
    reference on one host core beside it when oracle/_ref is present.  Latency, not throughput."""
    drop = os.path.join(ROOT, "apda-fft_b200")
    if drop not in sys.path:
        sys.path.insert(0, drop)
    import apda_fft_b200.synth as synth
    from metrics.fft_iterativa import start_fft
    from utils.get_peak_prominence import get_top_peaks_prominence
    axes = [synth.fleet_window(w, 1024).tolist() for w in range(3)]
    for ax in axes:
        get_top_peaks_prominence(start_fft(ax, 125.0), 125.0)
    reps = 50
    t0 = time.perf_counter()
    for _ in range(reps):
        peaks = [get_top_peaks_prominence(start_fft(ax, 125.0), 125.0) for ax in axes]
    ms = (time.perf_counter() - t0) / reps * 1e3
    out = {"workload": "3 axes x N=1024 fp64: start_fft + get_top_peaks_prominence through the drop-in modules (host lists "
                       "in, list of dicts out; H2D, kernels, D2H and the Python conversions inside the clock)",
           "ms_per_3axis_sensor": ms, "idx": [[p["idx"] for p in pk] for pk in peaks]}
    try:
        from oracle import ref_copy
        if ref_copy.available():
            ref = ref_copy.RefModules()
            t0 = time.perf_counter()
            want = [ref.get_top_peaks_prominence(ref.start_fft(ax, 125.0), 125.0) for ax in axes]
            out["reference_ms_per_3axis_sensor_1_core"] = (time.perf_counter() - t0) * 1e3
            out["equal_to_reference"] = bool(want == peaks)
    except Exception as exc:  # noqa: BLE001 - the reference copy is optional
        out["reference_note"] = f"{type(exc).__name__}: {exc}"
    return out


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
