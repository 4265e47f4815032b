#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that show what the shipped library uses (cuobjdump -sass on libapda_b200.so):
TMA (UTMALDG / UTMASTG), mbarriers (SYNCS), L2 prefetch requests (CCTL.E.PF2), constant-bank twiddles (LDCU.128), packed fp32 (FADD2 / FMUL2 / FFMA2), warp reductions (REDUX), fp64 pipe
(DFMA / DADD / DMUL), MUFU, 128-bit global accesses, and the absence of tensor-core / library code.

    python scripts/sass_counts.py > profiles/sass_r2.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "apda-fft_b200", "libapda_b200.so")
PATTERNS = ["UTMALDG", "UTMASTG", "SYNCS", "CCTL.E.PF2", "LDCU.128", "FADD2", "FMUL2", "FFMA2", "REDUX", "DFMA", "DADD", "DMUL", "MUFU",
            "LDG.E.128", "STG.E.128", "LDS.128", "STS.128", "BAR.SYNC", "HMMA", "UTCHMMA", "UTCQMMA"]


def strip_params(name: str) -> str:
    """Drop the trailing parameter list "(...)" of a demangled function, keeping template arguments such as (int)4096."""
    if not name.endswith(")"):
        return name
    depth = 0
    for i in range(len(name) - 1, -1, -1):
        depth += name[i] == ")"
        depth -= name[i] == "("
        if depth == 0:
            return name[:i]
    return name


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts = collections.OrderedDict()
    name = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["cu++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            name = strip_params(name)
            counts[name] = collections.Counter()
            continue
        if name is None:
            continue
        counts[name]["instructions"] += bool(re.search(r"/\*[0-9a-f]{4,5}\*/", line))
        for pat in PATTERNS:
            if re.search(r"\b" + re.escape(pat), line):
                counts[name][pat] += 1
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}: mnemonic counts per kernel (sm_100a)")
    print("# " + " ".join(["instr"] + PATTERNS))
    for fn, c in counts.items():
        cells = [str(c["instructions"])] + [str(c[p]) for p in PATTERNS]
        print(f"{fn}\n    " + " ".join(f"{p}={v}" for p, v in zip(["instr"] + PATTERNS, cells) if v != "0"))
    total = collections.Counter()
    for c in counts.values():
        total.update(c)
    print("# totals: " + " ".join(f"{p}={total[p]}" for p in ["instructions"] + PATTERNS))


if __name__ == "__main__":
    sys.exit(main())
