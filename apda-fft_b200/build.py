"""Build libapda_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python apda-fft_b200/build.py [--force]

One nvcc invocation per translation unit (in parallel), then one link.  The .so is git-ignored but travels with
the gpurun snapshot, so the GPU box never needs to compile.
"""
from __future__ import annotations

import hashlib
import os
import shlex
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libapda_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC,-O2,-ffp-contract=off"]
# extra nvcc flags for A/B builds (e.g. APDA_NVCC_EXTRA="-DAPDA_K1_MINB=8"); part of every unit's digest
EXTRA = shlex.split(os.environ.get("APDA_NVCC_EXTRA", ""))
LIB = os.environ.get("APDA_LIB_OUT", LIB)      # A/B builds: another output library ...
OBJ = os.environ.get("APDA_OBJ_DIR", OBJ)      # ... and its own object directory
# the device generator mirrors the host generator's separately rounded arithmetic
PER_FILE = {"synth.cu": ["-fmad=false"]}


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def headers_digest() -> bytes:
    """SHA-256 over every header a translation unit may include (all of them: the set is small)."""
    paths = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    paths.append(os.path.join(HERE, "..", "include", "apda_b200.h"))
    h = hashlib.sha256()
    for p in paths:
        with open(p, "rb") as fh:
            h.update(os.path.basename(p).encode() + b"\0" + fh.read() + b"\0")
    return h.digest()


def unit_digest(src: str, hdr: bytes) -> str:
    """What the object file depends on: the source text, every header, the compiler flags and the compiler itself.
    Content hashes, not mtimes: an object that travelled with a snapshot (or survived a checkout that reset the
    timestamps) is only reused when it was built from exactly these bytes."""
    h = hashlib.sha256(hdr)
    with open(os.path.join(CSRC, src), "rb") as fh:
        h.update(fh.read())
    h.update(" ".join([NVCC, *FLAGS, *PER_FILE.get(src, []), *EXTRA]).encode())
    return h.hexdigest()


def compile_one(src: str, force: bool, verbose: bool, hdr: bytes) -> str:
    obj = os.path.join(OBJ, src[:-3] + ".o")
    stamp = obj + ".sha256"
    want = unit_digest(src, hdr)
    if not force and os.path.exists(obj) and os.path.exists(stamp):
        with open(stamp) as fh:
            if fh.read().strip() == want:
                return obj
    cmd = [NVCC, *FLAGS, *PER_FILE.get(src, []), *EXTRA, "-c", os.path.join(CSRC, src), "-o", obj]
    cmd += ["-Xptxas", "-v"] if verbose else []
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}")
    with open(stamp, "w") as fh:
        fh.write(want + "\n")
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    hdr = headers_digest()
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as pool:
        objs = list(pool.map(lambda s: compile_one(s, force, verbose, hdr), sources()))
    # the library is stamped with the digests of the objects it was linked from
    link = hashlib.sha256()
    for o in objs:
        with open(o + ".sha256") as fh:
            link.update(fh.read().encode())
    want = link.hexdigest()
    stamp = LIB + ".sha256"
    fresh = False
    if not force and os.path.exists(LIB) and os.path.exists(stamp):
        with open(stamp) as fh:
            fresh = fh.read().strip() == want
    if not fresh:
        cmd = [NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
        subprocess.check_call(cmd)
        with open(stamp, "w") as fh:
            fh.write(want + "\n")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
