"""ctypes binding of ``libmcl_b200.so`` (the C ABI declared in ``include/mcl_b200.h``).

There is no CPU fallback: if the library is missing it is built with nvcc, and if that is not
possible importing a compute entry point fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

ABI_VERSION = 1

EXPORTS = (
    "mcl_abi_version", "mcl_last_error", "mcl_workspace_bytes", "mcl_run", "mcl_run_host",
    "mcl_device_peaks", "mcl_objective", "mcl_release_scratch", "mcl_debug_exp_draws", "mcl_objective_last_kernel_ms",
    "mcl_launch_count",
)


class HistSpec(C.Structure):
    _fields_ = [("axis", C.c_int32), ("n_bins", C.c_int32), ("n_groups", C.c_int32),
                ("reserved", C.c_int32), ("lo", C.c_double), ("hi", C.c_double)]


class RunArgs(C.Structure):
    _fields_ = [
        ("replicas", C.c_void_p), ("n_replicas", C.c_int32),
        ("segments", C.c_void_p), ("n_segments", C.c_int32),
        ("obs_time", C.c_void_p), ("n_obs", C.c_int32),
        ("max_steps", C.c_int32), ("mode", C.c_int32),
        ("seed", C.c_uint64), ("replica_id0", C.c_uint64),
        ("replay_u", C.c_void_p), ("replay_off", C.c_void_p),
        ("event", C.c_void_p), ("n_e", C.c_void_p), ("t", C.c_void_p),
        ("kind", C.c_void_p), ("e_idx", C.c_void_p), ("h_idx", C.c_void_p),
        ("steps_used", C.c_void_p), ("final_n_e", C.c_void_p), ("esteps", C.c_void_p),
        ("consumed", C.c_void_p), ("status", C.c_void_p), ("obs_n_e", C.c_void_p),
        ("hist", C.POINTER(HistSpec)), ("hist_group", C.c_void_p),
        ("hist_events", C.c_void_p), ("hist_occ", C.c_void_p), ("hist_occ_sq", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
        ("stream", C.c_void_p),
    ]


class Lab(C.Structure):
    _fields_ = [
        ("protocol", C.c_int32), ("n_rows", C.c_int32),
        ("rows", C.c_void_p), ("e_ratio_start", C.c_void_p), ("obs_begin", C.c_void_p),
        ("obs_time", C.c_void_p), ("target", C.c_void_p),
        ("N_e", C.c_double), ("boundary_factor", C.c_double), ("D", C.c_double), ("k_b", C.c_double),
        ("max_steps", C.c_int32), ("flags", C.c_int32),
    ]


class Peaks(C.Structure):
    _fields_ = [("mufu_gops", C.c_double), ("ffma_gops", C.c_double), ("imad_gops", C.c_double),
                ("lop3_gops", C.c_double), ("sm_clock_mhz", C.c_double),
                ("n_sm", C.c_int32), ("reserved", C.c_int32)]


_lib = None


class NativeError(RuntimeError):
    pass


def lib_path() -> str:
    return _build.LIB_PATH


def load():
    """Load (building first if needed) the shared library.  Raises if it cannot be had."""
    global _lib
    if _lib is not None:
        return _lib
    override = os.environ.get("MCL_B200_LIB")                   # override: A/B builds while tuning
    if override:
        if not os.path.isfile(override):
            raise NativeError(f"MCL_B200_LIB names a missing file: {override}")
        path = override
    else:
        path = _build.LIB_PATH
        # build() recompiles whatever is older than its sources (mtime), so an edited kernel is never
        # silently tested against a stale library; without nvcc an existing library is used as is
        try:
            _build.build()
        except Exception as exc:  # noqa: BLE001
            if not os.path.isfile(path):
                raise NativeError(
                    f"libmcl_b200.so is missing and could not be built ({exc}); "
                    "the kinetics path has no CPU fallback") from exc
            if _build.have_nvcc():
                raise NativeError(f"libmcl_b200.so is stale and the rebuild failed: {exc}") from exc
    L = C.CDLL(path)
    L.mcl_abi_version.restype = C.c_int
    L.mcl_last_error.restype = C.c_char_p
    L.mcl_workspace_bytes.restype = C.c_size_t
    L.mcl_workspace_bytes.argtypes = [C.POINTER(RunArgs)]
    L.mcl_run.restype = C.c_int
    L.mcl_run.argtypes = [C.POINTER(RunArgs)]
    L.mcl_run_host.restype = C.c_int
    L.mcl_run_host.argtypes = [C.POINTER(RunArgs)]
    L.mcl_launch_count.restype = C.c_int64
    L.mcl_launch_count.argtypes = []
    L.mcl_device_peaks.restype = C.c_int
    L.mcl_device_peaks.argtypes = [C.POINTER(Peaks)]
    if hasattr(L, "mcl_objective"):
        L.mcl_objective.restype = C.c_int
        L.mcl_objective.argtypes = [C.c_void_p, C.c_int32, C.POINTER(Lab), C.c_uint64, C.c_uint64,
                                    C.c_void_p, C.c_void_p, C.c_void_p]
    if hasattr(L, "mcl_objective_last_kernel_ms"):       # (older A/B builds named by MCL_B200_LIB may lack the newer hooks)
        L.mcl_objective_last_kernel_ms.restype = C.c_float
    if hasattr(L, "mcl_debug_exp_draws"):
        L.mcl_debug_exp_draws.restype = C.c_int
        L.mcl_debug_exp_draws.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    if L.mcl_abi_version() != ABI_VERSION:
        raise NativeError(f"libmcl_b200.so ABI {L.mcl_abi_version()} != expected {ABI_VERSION}")
    _lib = L
    return L


def last_error() -> str:
    return (load().mcl_last_error() or b"").decode("utf-8", "replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise NativeError(f"{what} failed with status {rc}: {last_error()}")
