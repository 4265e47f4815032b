#!/usr/bin/env python
"""Per-source-line profile of one kernel from an ncu report: joins the SASS rows of `ncu --page source --csv` with the
line table of the matching cubin (nvdisasm -g), and sums instructions executed and stall samples per source line.

    python scripts/ncu_lines2.py <report.ncu-rep> <kernel regex> <cubin> <mangled-name substring> [top]
"""
import collections
import csv
import io
import re
import subprocess
import sys


def main():
    rep, kregex, cubin, mangled = sys.argv[1:5]
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kregex}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    # the report may hold several kernels: take the first block
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    body = []
    for r in rows[hdr_i + 1:]:
        if not r or r[0] == "Kernel Name":
            break
        body.append(r)
    col = {h: i for i, h in enumerate(hdr)}
    base = int(body[0][col["Address"]], 16)
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    # nvdisasm: "//## File "x", line N" markers, then "        /*0970*/   OPCODE ..." inside ".text.<mangled>" sections
    line_of = {}
    cur = ("?", 0)
    active = False
    for ln in dis.splitlines():
        if ln.startswith("//--------------------- .text."):
            active = mangled in ln
            continue
        if not active:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/", ln)
        if m:
            line_of[int(m.group(1), 16)] = cur
    per = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    total_i = total_s = 0
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for r in body:
        off = int(r[col["Address"]], 16) - base
        key = line_of.get(off, ("?", 0))
        inst = int(r[col["Instructions Executed"]] or 0)
        samp = int(r[col["# Samples"]] or 0)
        per[key][0] += inst
        per[key][1] += samp
        for h in stall_cols:
            v = int(r[col[h]] or 0)
            if v:
                per[key][2][h] += v
        total_i += inst
        total_s += samp
    print(f"kernel {kregex}: {total_i} warp instructions, {total_s} samples")
    for key, (inst, samp, st) in sorted(per.items(), key=lambda kv: -kv[1][1])[:top]:
        tops = ", ".join(f"{k[6:]}={v}" for k, v in st.most_common(3))
        print(f"{key[0]}:{key[1]:<5d} inst {100 * inst / total_i:5.1f}%  samples {100 * samp / max(total_s, 1):5.1f}%  {tops}")


if __name__ == "__main__":
    main()
