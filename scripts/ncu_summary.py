#!/usr/bin/env python
"""Summarise an ncu report (ncu --set full) into profiles/: one markdown table of the metrics the roofline numbers
come from, plus profiles/traffic.json (DRAM bytes per launch of each kernel, consumed by bench.py's roofline.traffic).

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/ncu_r1 [--windows 200000] [--w <name regex>=<windows> ...]

--w gives the windows per launch of the kernels whose name matches the regex (reports that hold launches of different
batch sizes); --windows is the default for the others.
"""
import csv
import io
import json
import os
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm % of peak"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs) blocks"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem) blocks"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (expected 0)"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe busy %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma pipe busy %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "alu pipe busy %"),
]


def main():
    rep, out_prefix = sys.argv[1], sys.argv[2]
    default_windows = int(sys.argv[sys.argv.index("--windows") + 1]) if "--windows" in sys.argv else None
    import re
    wmap = []
    for i, a in enumerate(sys.argv):
        if a == "--w":
            rx, _, cnt = sys.argv[i + 1].rpartition("=")
            wmap.append((re.compile(rx), int(cnt)))
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = [f"# ncu summary of `{os.path.basename(rep)}`", "",
             "`ncu --set full --clock-control none --import-source on` on B200 (sm_100a); one row block per captured launch.",
             "Per-launch times under ncu are cold-cache and serialised: compare shares and byte counts, not absolute times.", ""]
    traffic = {}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        windows = next((cnt for rx, cnt in wmap if rx.search(name)), default_windows)
        lines += [f"## `{name[:140]}`", "", "| metric | value | unit |", "|---|---|---|"]
        vals = {}
        for key, label in METRICS:
            if key in hdr:
                i = hdr.index(key)
                vals[key] = r[i]
                lines.append(f"| {label} (`{key}`) | {r[i]} | {units[i]} |")
        stalls = sorted(((float(r[i] or 0), h) for i, h in enumerate(hdr)
                         if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")),
                        reverse=True)[:5]
        lines.append("| top stall reasons (warps per issue) | " + ", ".join(
            f"{h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]} {v:.2f}" for v, h in stalls) + " | |")
        if windows:
            lines.append(f"| windows per launch | {windows} | |")
            try:
                lines.append(f"| warp instructions per window | {float(vals['smsp__inst_executed.sum']) / windows:.0f} | |")
            except Exception:  # noqa: BLE001
                pass
        try:
            def to_bytes(v, u):
                v = float(v.replace(",", ""))
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            rd = to_bytes(vals["dram__bytes_read.sum"], units[hdr.index("dram__bytes_read.sum")])
            wr = to_bytes(vals["dram__bytes_write.sum"], units[hdr.index("dram__bytes_write.sum")])
            entry = {"dram_bytes_per_launch": rd + wr, "read": rd, "write": wr, "windows": windows}
            if windows:
                entry["dram_bytes_per_window"] = (rd + wr) / windows
            m = re.search(r"(fft_f32_fast|fft_f64_fast|peaks_f32_fast|peaks_f64_fast|fft_smem|peaks)_kernel<\(?(?:int\))?(\d+)", name)
            if m and m.group(1) == "fft_f64_fast":      # templated on log2(N)
                key = f"k1_f64_n{1 << int(m.group(2))}"
            elif m and m.group(1).startswith("fft"):
                key = f"k1_f32_n{m.group(2)}"
            elif m and m.group(1) in ("peaks_f32_fast", "peaks_f64_fast"):
                key = f"k3_{m.group(1)[6:9]}_n{2 * int(m.group(2))}"
            else:
                key = name[:60]
            traffic.setdefault(key, entry)
            lines.append(f"| DRAM traffic per launch | {(rd + wr) / 1e9:.3f} | GB |")
            if windows:
                lines.append(f"| DRAM traffic per window | {(rd + wr) / windows:.0f} | B |")
        except Exception as exc:  # noqa: BLE001
            lines.append(f"| traffic | n/a ({exc}) | |")
        lines.append("")
    with open(out_prefix + "_summary.md", "w") as fh:
        fh.write("\n".join(lines) + "\n")
    tpath = os.path.join(os.path.dirname(out_prefix), "traffic.json")
    try:
        merged = json.load(open(tpath))
    except Exception:  # noqa: BLE001
        merged = {}
    merged.update(traffic)
    with open(tpath, "w") as fh:
        json.dump(merged, fh, indent=1)
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main()
