"""Synthetic workloads of the shapes BASELINE.json names, as replica tables.

C1  default conf/config_fp.yaml (TL12 + basicTL12): 4 heating rates x 2 sims, N_e = 2000
C2  isothermal hold then optical readout: R replicas x N_e = 10^4, ensemble L(t) on log-time bins
C3  dose response: irradiation from empty to 10 dose points, then TL readout, N_e = 2000
C4  Optimizer inner loop: S Sobol candidates x 11 lab rows (tl_clbr), n_e <= 100
C5  large ensemble: R replicas x C1 geometry at 1 degC/s (10^8 electrons = 5*10^4 replicas)
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

from .config import compose, initialize_runs
from .engine import AXIS_TEMP, AXIS_TIME_LOG, HistSpec
from .replicas import PROTO_SIMULATE, REPLICA_DTYPE, SEGMENT_DTYPE, fill_replica, simulate_tables


def c1():
    cfg = compose()
    runs = initialize_runs(cfg)
    sims = int(runs[0]["exp_type_fp"]["sims"])
    reps, segs = simulate_tables(runs, sims)
    return dict(name="C1 default config_fp.yaml (TL12+basicTL12, 4 T_rate x 2 sims, N_e=2000)",
                replicas=reps, segments=segs, max_steps=int(runs[0]["exp_type_fp"]["steps"]),
                hist=None, hist_group=None)


#: C2 schedule: hold at 250 degC for 1000 s, then optical readout at 50 degC, A_opt = 100 1/s, for 10^4 s
C2_HOLD = dict(T_start=250.0, T_rate=0.0, duration=1.0e3, dose_rate=0.0, dt_cap=1e20, A_opt=0.0)
C2_OSL = dict(T_start=50.0, T_rate=0.0, duration=1.0e4, dose_rate=0.0, dt_cap=1e20, A_opt=100.0)


def _segments(*legs) -> np.ndarray:
    segs = np.zeros(len(legs), dtype=SEGMENT_DTYPE)
    for i, leg in enumerate(legs):
        for k, v in leg.items():
            segs[i][k] = v
    return segs


def c2(n_replicas: int = 10_000, n_e: int = 10_000, n_bins: int = 1000, physics_overrides=()):
    """`physics_overrides` (e.g. two distinct tunnelling channels: ["physics_fp.E_loc_2=1.0", "physics_fp.Retrap=0.3"]) are for
    side measurements only; the BASELINE workload uses the shipped basicTL12 physics."""
    cfg = compose(overrides=[f"exp_type_fp.N_e={n_e}", f"exp_type_fp.holes={n_e}",
                             "exp_type_fp.T_rate=[0]", "exp_type_fp.duration=[1000]", "exp_type_fp.sims=1", *physics_overrides])
    run = initialize_runs(cfg)[0]
    rec = np.zeros(1, dtype=REPLICA_DTYPE)
    fill_replica(rec[0], run["exp_type_fp"], run["physics_fp"], 1.0, PROTO_SIMULATE)
    rec["seg_begin"], rec["seg_count"] = 0, 2
    reps = np.repeat(rec, n_replicas)
    segs = _segments(C2_HOLD, C2_OSL)
    hist = HistSpec(axis=AXIS_TIME_LOG, n_bins=n_bins, lo=1e-3, hi=2e4, n_groups=2)   # one row per leg
    return dict(name=f"C2 isothermal hold 250C/1000s + OSL readout 50C/A=100/1e4s, {n_replicas} replicas x {n_e} electrons",
                replicas=reps, segments=segs, max_steps=4 * n_e, hist=hist, hist_group=None)


def c5(n_replicas: int = 50_000, n_bins: int = 800):
    cfg = compose(overrides=["exp_type_fp.T_rate=[1]", "exp_type_fp.duration=[800]", "exp_type_fp.sims=1"])
    run = initialize_runs(cfg)
    reps1, segs = simulate_tables(run, 1)
    reps = np.repeat(reps1, n_replicas)
    hist = HistSpec(axis=AXIS_TEMP, n_bins=n_bins, lo=0.0, hi=800.0, n_groups=1)
    return dict(name=f"C5 TL ramp 1 degC/s, {n_replicas} replicas x 2000 electrons",
                replicas=reps, segments=segs, max_steps=20000, hist=hist, hist_group=None)


def c3(replicas_per_dose: int = 256, n_bins: int = 800):
    """10 irradiation times (as in lab ISO experiment 0, log-spaced 543..16287 s) at 15 degC and
    D = 0.092 Gy/s from empty traps, then a TL ramp 0 -> 800 degC at 1 degC/s; lab_TL physics."""
    cfg = compose(overrides=["physics_fp=lab_TL", "exp_type_fp.N_e=2000", "exp_type_fp.holes=2000",
                             "exp_type_fp.e_ratio_start=0", "exp_type_fp.T_rate=[1]",
                             "exp_type_fp.duration=[800]", "exp_type_fp.sims=1"])
    run = initialize_runs(cfg)[0]
    doses = np.geomspace(543.0, 16287.0, 10)
    rec = np.zeros(1, dtype=REPLICA_DTYPE)
    fill_replica(rec[0], run["exp_type_fp"], run["physics_fp"], 0.0, PROTO_SIMULATE)
    legs = []
    reps = np.repeat(rec, 10 * replicas_per_dose)
    group = np.zeros(len(reps), dtype=np.int32)
    for d, dur in enumerate(doses):
        legs.append(dict(T_start=15.0, T_rate=0.0, duration=float(dur), dose_rate=0.092, dt_cap=1e20, A_opt=0.0))
        legs.append(dict(T_start=0.0, T_rate=1.0, duration=800.0, dose_rate=0.0, dt_cap=1.0, A_opt=0.0))
        sl = slice(d * replicas_per_dose, (d + 1) * replicas_per_dose)
        reps["seg_begin"][sl], reps["seg_count"][sl] = 2 * d, 2
        group[sl] = 2 * d                       # row of the replica's first leg; leg sg writes row 2 d + sg
    hist = HistSpec(axis=AXIS_TEMP, n_bins=n_bins, lo=0.0, hi=800.0, n_groups=20)
    return dict(name=f"C3 dose response: 10 doses x {replicas_per_dose} replicas, irradiation then TL, N_e=2000",
                replicas=reps, segments=_segments(*legs), max_steps=40000, hist=hist, hist_group=group)


def c4_candidates(n: int = 4096, seed: int = 4) -> np.ndarray:
    """Sobol points in the Optimizer's DEFAULT_BOUNDS, shape [10, n] (parameter-major)."""
    from scipy.stats import qmc
    from .optimizer import DEFAULT_BOUNDS
    b = np.asarray(DEFAULT_BOUNDS, dtype=float)
    pts = qmc.Sobol(d=10, seed=seed).random(n)
    return np.ascontiguousarray((b[:, 0] + pts * (b[:, 1] - b[:, 0])).T)
