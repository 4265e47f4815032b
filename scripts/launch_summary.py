#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) into a per-kernel share table.

    python scripts/launch_summary.py gpurun_out/launches.csv profiles/launches_rNN_summary.md "command that was profiled"
"""
import collections
import csv
import re
import sys

src, dst, cmd = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "bench.py"
rows = [r for r in csv.reader(open(src, errors="replace")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows:
    if r is hdr or len(r) <= iv or r[ik] == "Kernel Name":
        continue
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    unit = r[iu]
    ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    name = re.sub(r"\(.*", "", r[ik]).strip()
    name = re.sub(r"\((int|bool)\)", "", name)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ms
total = sum(v[1] for v in agg.values()) or 1.0
lines = [f"# ncu launch list of `{cmd}`, summarised", "",
         "`ncu --metrics gpu__time_duration.sum --clock-control none`; raw rows beside this file. Times under ncu are "
         "cold-cache and serialised: shares, not absolutes.", "", "| kernel | launches | total ms | share |", "|---|---|---|---|"]
for name, (cnt, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"| `{name[:110]}` | {cnt} | {ms:.3f} | {100 * ms / total:.1f}% |")
open(dst, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
