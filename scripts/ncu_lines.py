#!/usr/bin/env python
"""Per-source-line instruction / stall-sample totals of one kernel from an ncu report (needs -lineinfo + --import-source on).

    python scripts/ncu_lines.py report.ncu-rep build/obj.o 'kernelILi2048ELb1' [top]

ncu's SASS page is matched by instruction order with nvdisasm -g of the same object (line info).
"""
import csv, io, os, re, subprocess, sys, tempfile, collections

rep, obj, pat = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
# several kernels may be in the report: split on "Kernel Name" rows
blocks, cur = [], None
for row in csv.reader(io.StringIO(raw)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(row)
want = re.sub(r"ILi(\d+)ELb(\d)", "", pat)
blk = None
for b in blocks:
    nm = b["name"]
    m = re.search(r"ILi(\d+)ELb(\d)", pat)
    if want.split("IL")[0] in nm.replace("::", "") or True:
        if m and (f"(int){m.group(1)}" in nm and f"(bool){m.group(2)}" in nm) and pat.split("IL")[0].split("kernel")[0] in nm:
            blk = b
            break
if blk is None:
    blk = blocks[0]
hdr = blk["rows"][0]
ix, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
sass = [(r[ix].strip(), int(r[ie] or 0), int(r[isamp] or 0)) for r in blk["rows"][1:] if len(r) > isamp]

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
sect, on, line, lines = [], False, "?", []
for l in dis.splitlines():
    if l.startswith("//--------------------- .text."):
        on = pat in l
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        line = f"{os.path.basename(m.group(1))}:{m.group(2)}"
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", l)
    if m:
        lines.append((line, m.group(1).strip()))
print(f"kernel {blk['name'][:90]}\n  ncu sass rows {len(sass)}, nvdisasm instrs {len(lines)}")
n = min(len(sass), len(lines))
agg = collections.defaultdict(lambda: [0, 0])
for (src, ex, smp), (ln, ins) in zip(sass[:n], lines[:n]):
    agg[ln][0] += ex
    agg[ln][1] += smp
tot_e = sum(v[0] for v in agg.values()) or 1
tot_s = sum(v[1] for v in agg.values()) or 1
print(f"  total warp-instr {tot_e}, samples {tot_s}")
for ln, (e, s) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{ln:32s} instr {100*e/tot_e:5.1f}%  samples {100*s/tot_s:5.1f}%")
