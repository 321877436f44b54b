"""CPU-side checks of the boundary: the library builds, loads and exports what the header declares."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "mcl_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mcl_[a-z_0-9]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    from mcluminescence_b200 import _native
    L = _native.load()
    names = declared_functions()
    assert "mcl_run" in names and "mcl_run_host" in names and "mcl_objective" in names
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/mcl_b200.h but not exported"
    assert L.mcl_abi_version() == _native.ABI_VERSION


def test_struct_layouts_match_header():
    from mcluminescence_b200 import _native
    from mcluminescence_b200.replicas import REPLICA_DTYPE, SEGMENT_DTYPE
    assert SEGMENT_DTYPE.itemsize == 48 and REPLICA_DTYPE.itemsize == 120
    assert C.sizeof(_native.HistSpec) == 32
    assert C.sizeof(_native.Peaks) == 48
    # pointer + int32 pairs are padded to 16 bytes each in mcl_run_args
    assert _native.RunArgs.n_replicas.offset == 8 and _native.RunArgs.segments.offset == 16
    assert _native.RunArgs.seed.offset == 56 and _native.RunArgs.replay_u.offset == 72


def test_workspace_query_and_argument_validation_need_no_gpu(golden):
    from mcluminescence_b200 import _native
    from tests import helpers
    L = _native.load()
    reps, segs, steps, _, _ = helpers.sim_tables_for(golden.meta("sim_kat1"))
    a = _native.RunArgs()
    a.replicas, a.n_replicas = reps.ctypes.data, len(reps)
    a.segments, a.n_segments = segs.ctypes.data, len(segs)
    a.max_steps, a.mode = steps, 0
    assert L.mcl_workspace_bytes(C.byref(a)) > 0
    a.mode = 1
    assert L.mcl_workspace_bytes(C.byref(a)) > 0
    a.mode = 7
    assert L.mcl_workspace_bytes(C.byref(a)) == 0 and b"mode" in L.mcl_last_error()
    bad = reps.copy()
    bad["seg_begin"] = 5
    a.mode, a.replicas = 0, bad.ctypes.data
    assert L.mcl_workspace_bytes(C.byref(a)) == 0 and b"segment range" in L.mcl_last_error()


def test_product_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from mcluminescence_b200 import _native, engine
    from mcluminescence_b200.replicas import REPLICA_DTYPE, SEGMENT_DTYPE
    with pytest.raises(_native.NativeError, match="no CPU fallback"):
        engine.run_replicas(np.zeros(1, REPLICA_DTYPE), np.zeros(1, SEGMENT_DTYPE), 10)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "mcluminescence_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "mcl_oracle" not in text and "from oracle" not in text and "import oracle" not in text, f
