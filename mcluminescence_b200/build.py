"""Build libmcl_b200.so (hand-written sm_100a CUDA + the C ABI) in-tree with nvcc.

    python -m mcluminescence_b200.build [--force]

nvcc cross-compiles without a GPU.  The library lands in ``mcluminescence_b200/_lib/`` so that it
travels with the source tree to the GPU box; it is git-ignored.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.abspath(os.path.join(HERE, "..", "include"))
LIB_DIR = os.path.join(HERE, "_lib")
LIB_PATH = os.path.join(LIB_DIR, "libmcl_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "--extended-lambda", "-Xcompiler", "-fPIC",
          "-I", INCLUDE, "-I", CSRC]
# per-file extras: the replay kernel must never fuse a multiply into an add (NumPy does not)
EXTRA = {"mcl_replay.cu": ["-fmad=false"]}
SOURCES = ["mcl_abi.cu", "mcl_replay.cu", "mcl_philox.cu", "mcl_peaks.cu", "mcl_objective.cu"]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.isfile(exe):
        raise RuntimeError("nvcc not found; libmcl_b200.so cannot be built")
    return exe


def _stale(target: str, deps) -> bool:
    if not os.path.isfile(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIB_DIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(INCLUDE, "mcl_b200.h"))
    objs = []
    for src in SOURCES:
        sp = os.path.join(CSRC, src)
        if not os.path.isfile(sp):
            continue
        obj = os.path.join(LIB_DIR, src.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [sp] + headers):
            cmd = [nvcc()] + ARCH + COMMON + EXTRA.get(src, []) + (["-Xptxas", "-v"] if verbose else []) \
                + ["-c", sp, "-o", obj]
            subprocess.run(cmd, check=True)
    if force or _stale(LIB_PATH, objs):
        subprocess.run([nvcc()] + ARCH + ["-shared", "-o", LIB_PATH] + objs, check=True)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
