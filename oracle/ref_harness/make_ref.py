"""Recipe for ``oracle/_ref``: the UNMODIFIED reference modules of the hot path, so that the NumPy reference itself
can be timed on the GPU box's host cores next to the C port (bench.py `cpu_baseline.numpy_reference`).

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  ``/root/reference`` exists only in the build container; ``oracle/_ref/`` is
git-ignored (never enters history) but travels to the GPU box with the tree, like the built libraries.  The five files
are copied byte for byte (MIT licence, copied along) and a manifest of their SHA-256 digests is written; nothing in
the product imports them.  Run by ``__graft_entry__.build()`` when the reference tree is present.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("MCL_REFERENCE_ROOT", "/root/reference")
OUT = os.path.abspath(os.path.join(HERE, "..", "_ref"))
FILES = ["engine.py", "tl_trap_lab.py", "simulate.py", "optimizer.py", "paths.py"]


def make() -> bool:
    src = os.path.join(REF_ROOT, "src", "class")
    if not all(os.path.isfile(os.path.join(src, f)) for f in FILES):
        return False
    dst = os.path.join(OUT, "src_class")
    os.makedirs(dst, exist_ok=True)
    manifest = {}
    for f in FILES:
        shutil.copyfile(os.path.join(src, f), os.path.join(dst, f))
        with open(os.path.join(dst, f), "rb") as fh:
            manifest[f] = hashlib.sha256(fh.read()).hexdigest()
    lic = os.path.join(REF_ROOT, "LICENSE")
    if os.path.isfile(lic):
        shutil.copyfile(lic, os.path.join(OUT, "LICENSE"))
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as fh:
        json.dump({"source": "HarrisNH/MCLuminescence src/class (unmodified)", "sha256": manifest}, fh, indent=1)
    return True


if __name__ == "__main__":
    print("oracle/_ref written" if make() else "reference tree not found; oracle/_ref not written")
