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
SOURCES = ["mcl_abi.cu", "mcl_replay.cu", "mcl_philox.cu", "mcl_smallbox.cu", "mcl_peaks.cu", "mcl_objective.cu"]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.isfile(exe):
        raise RuntimeError("nvcc not found; libmcl_b200.so cannot be built")
    return exe


def have_nvcc() -> bool:
    try:
        nvcc()
        return True
    except RuntimeError:
        return False


def _digest(paths, extra=()) -> str:
    """Content fingerprint of a compilation unit: sources, headers and flags.  (mtimes do not survive the
    copy to the GPU box, and a stale library must never be loaded silently.)"""
    import hashlib
    h = hashlib.sha256()
    for p in paths:
        with open(p, "rb") as fh:
            h.update(fh.read())
        h.update(b"\0")
    h.update("\0".join(extra).encode())
    return h.hexdigest()


def _stamp_ok(target: str, digest: str) -> bool:
    try:
        with open(target + ".stamp") as fh:
            return os.path.isfile(target) and fh.read().strip() == digest
    except OSError:
        return False


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile what is out of date (by content fingerprint) and link.  Safe to call from several processes
    at once (one rank per GPU): a file lock serialises them and the later ones find everything fresh."""
    import fcntl
    os.makedirs(LIB_DIR, exist_ok=True)
    headers = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    headers.append(os.path.join(INCLUDE, "mcl_b200.h"))
    with open(os.path.join(LIB_DIR, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        objs, stamps = [], []
        for src in SOURCES:
            sp = os.path.join(CSRC, src)
            if not os.path.isfile(sp):
                continue
            obj = os.path.join(LIB_DIR, src.replace(".cu", ".o"))
            flags = ARCH + COMMON + EXTRA.get(src, [])
            dg = _digest([sp] + headers, flags)
            objs.append(obj)
            stamps.append(dg)
            if force or not _stamp_ok(obj, dg):
                cmd = [nvcc()] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", sp, "-o", obj]
                subprocess.run(cmd, check=True)
                with open(obj + ".stamp", "w") as fh:
                    fh.write(dg)
        lib_dg = _digest([], stamps)
        if force or not _stamp_ok(LIB_PATH, lib_dg):
            subprocess.run([nvcc()] + ARCH + ["-shared", "-o", LIB_PATH] + objs, check=True)
            with open(LIB_PATH + ".stamp", "w") as fh:
                fh.write(lib_dg)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
