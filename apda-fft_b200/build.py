"""Build libapda_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python apda-fft_b200/build.py [--force]

One nvcc invocation per translation unit (in parallel), then one link.  The .so is git-ignored but travels with
the gpurun snapshot, so the GPU box never needs to compile.
"""
from __future__ import annotations

import os
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
# the device generator mirrors the host generator's separately rounded arithmetic
PER_FILE = {"synth.cu": ["-fmad=false"]}


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def headers_mtime():
    paths = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    paths.append(os.path.join(HERE, "..", "include", "apda_b200.h"))
    return max(os.path.getmtime(p) for p in paths)


def compile_one(src: str, force: bool, verbose: bool) -> str:
    obj = os.path.join(OBJ, src[:-3] + ".o")
    spath = os.path.join(CSRC, src)
    if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(spath), headers_mtime()):
        return obj
    cmd = [NVCC, *FLAGS, *PER_FILE.get(src, []), "-c", spath, "-o", obj] + (["-Xptxas", "-v"] if verbose else [])
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}")
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as pool:
        objs = list(pool.map(lambda s: compile_one(s, force, verbose), sources()))
    if force or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
