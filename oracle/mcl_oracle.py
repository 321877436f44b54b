"""ctypes front-end of the CPU oracle (``oracle/mcl_oracle.c``).  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs import this.
It accepts the same replica / segment tables the product builds (``mcluminescence_b200.replicas``)
because ``mclo_replica`` / ``mclo_segment`` share their layout with the C ABI structs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libmcl_oracle.so")
_lib = None

SEGMENT_DTYPE = np.dtype(
    [("T_start", "<f8"), ("T_rate", "<f8"), ("duration", "<f8"), ("dose_rate", "<f8"),
     ("dt_cap", "<f8"), ("A_opt", "<f8")]
)
REPLICA_DTYPE = np.dtype(
    [("alpha", "<f8"), ("b", "<f8"), ("s", "<f8"), ("E_cb", "<f8"), ("E_loc_1", "<f8"),
     ("E_loc_2", "<f8"), ("D0", "<f8"), ("Retrap", "<f8"), ("k_b", "<f8"), ("side", "<f8"),
     ("boundary_factor", "<f8"),
     ("N_e", "<i4"), ("n_e0", "<i4"), ("n_h0", "<i4"), ("protocol", "<i4"),
     ("seg_begin", "<i4"), ("seg_count", "<i4"), ("obs_begin", "<i4"), ("obs_count", "<i4")]
)

ERRORS = {-1: "steps", -2: "noholes", -3: "stream", -4: "alloc", -5: "noevent"}


class _Out(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "event", "n_e", "t", "kind", "e_idx", "h_idx", "steps_used", "final_n_e", "consumed",
        "esteps", "obs_n_e", "obs_t", "status")]


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (seconds).  Building the checker is not using it."""
    src = os.path.join(_HERE, "mcl_oracle.c")
    if force or not os.path.isfile(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "--no-print-directory"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.mclo_rng_sizeof.restype = C.c_size_t
        _lib.mclo_rng_consumed.restype = C.c_int64
        _lib.mclo_run_sequential.restype = C.c_int
        _lib.mclo_run_parallel.restype = C.c_int
        _lib.mclo_max_threads.restype = C.c_int
    return _lib


class Rng:
    """NumPy-legacy MT19937 stream (``np.random.seed(seed)``) or an external uniform buffer."""

    def __init__(self, seed: Optional[int] = None, external: Optional[np.ndarray] = None):
        L = lib()
        self._buf = C.create_string_buffer(L.mclo_rng_sizeof())
        self._ext = None
        if external is not None:
            self._ext = np.ascontiguousarray(external, dtype=np.float64)
            L.mclo_rng_external(self._buf, self._ext.ctypes.data_as(C.c_void_p),
                                C.c_int64(self._ext.size))
        else:
            L.mclo_rng_seed(self._buf, C.c_uint32(int(seed) & 0xFFFFFFFF))

    def uniforms(self, n: int) -> np.ndarray:
        out = np.empty(n, dtype=np.float64)
        lib().mclo_rng_fill(self._buf, out.ctypes.data_as(C.c_void_p), C.c_int64(n))
        return out

    @property
    def consumed(self) -> int:
        return int(lib().mclo_rng_consumed(self._buf))


@dataclass
class Result:
    event: np.ndarray
    n_e: np.ndarray
    t: np.ndarray
    kind: np.ndarray
    e_idx: np.ndarray
    h_idx: np.ndarray
    steps_used: np.ndarray
    final_n_e: np.ndarray
    consumed: np.ndarray
    esteps: np.ndarray
    obs_n_e: np.ndarray
    obs_t: np.ndarray
    status: np.ndarray
    rc: int


def _tables(replicas, segments, obs_time):
    reps = np.ascontiguousarray(replicas).view(np.uint8).view(REPLICA_DTYPE) \
        if replicas.dtype.itemsize == REPLICA_DTYPE.itemsize else None
    if reps is None:
        raise ValueError("replica table has the wrong record size")
    segs = np.ascontiguousarray(segments).view(np.uint8).view(SEGMENT_DTYPE)
    obs = np.ascontiguousarray(obs_time if obs_time is not None else np.zeros(0), dtype=np.float64)
    return reps, segs, obs


def run(replicas, segments, max_steps: int, *, rng: Optional[Rng] = None, seed: Optional[int] = None,
        obs_time=None, parallel: bool = False, threads: int = 0, trace: bool = True) -> Result:
    """Run replicas through the oracle.

    ``parallel=False``: one continuing stream (``rng`` or ``Rng(seed)``), replicas in order --
    the reference's semantics.  ``parallel=True``: replica r gets stream ``seed + r`` and
    replicas are spread over ``threads`` host threads (CPU-baseline leg).
    """
    L = lib()
    reps, segs, obs = _tables(replicas, segments, obs_time)
    R = reps.shape[0]
    n_obs = max(int(obs.size), 1)
    shape = (R, max_steps) if trace else (0, 0)
    r = Result(
        event=np.zeros(shape, np.int32), n_e=np.zeros(shape, np.int32), t=np.zeros(shape, np.float64),
        kind=np.zeros(shape, np.int32), e_idx=np.full(shape, -1, np.int32),
        h_idx=np.full(shape, -1, np.int32),
        steps_used=np.zeros(R, np.int32), final_n_e=np.zeros(R, np.int32),
        consumed=np.zeros(R, np.int64), esteps=np.zeros(R, np.int64),
        obs_n_e=np.full(n_obs, -1, np.int32), obs_t=np.zeros(n_obs, np.float64),
        status=np.zeros(R, np.int32), rc=0)
    out = _Out()
    for name in ("event", "n_e", "t", "kind", "e_idx", "h_idx"):
        setattr(out, name, getattr(r, name).ctypes.data if trace else None)
    for name in ("steps_used", "final_n_e", "consumed", "esteps", "obs_n_e", "obs_t", "status"):
        setattr(out, name, getattr(r, name).ctypes.data)
    rp = reps.ctypes.data_as(C.c_void_p)
    sp = segs.ctypes.data_as(C.c_void_p)
    op = obs.ctypes.data_as(C.c_void_p)
    if parallel:
        nt = threads if threads > 0 else L.mclo_max_threads()
        r.rc = L.mclo_run_parallel(rp, C.c_int(R), sp, op, C.c_int(max_steps),
                                   C.c_uint32(int(seed or 0) & 0xFFFFFFFF), C.c_int(nt), C.byref(out))
    else:
        if rng is None:
            rng = Rng(seed if seed is not None else 0)
        r.rc = L.mclo_run_sequential(rp, C.c_int(R), sp, op, C.c_int(max_steps), rng._buf,
                                     C.byref(out))
    return r


def max_threads() -> int:
    return int(lib().mclo_max_threads())
