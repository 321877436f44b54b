"""Host-side replica tables: what ``TLTrapSim.__init__`` derives, laid out for the C ABI.

``REPLICA_DTYPE`` / ``SEGMENT_DTYPE`` mirror ``mcl_replica`` / ``mcl_segment`` in
``include/mcl_b200.h`` byte for byte.  The scalar geometry is computed with the reference's own
Python expressions (``src/class/tl_trap_lab.py:33-39``, ``src/class/engine.py:124-128``) so that
``int()`` truncation and float rounding are identical: e.g. ``int(2000 * 1.2**3) == 3455``.
"""
from __future__ import annotations

from typing import Any, List, Mapping, Optional, Sequence, Tuple

import numpy as np

from .config import physics_record

PROTO_SIMULATE = 0
PROTO_TL_LAB = 1
PROTO_ISO_LAB = 2
PROTO_TL_LEGACY = 3          # pre-refactor TL loop (reference src/est_params/functions.py:270-360), see legacy_box_geometry

MODE_PHILOX = 0
MODE_REPLAY = 1

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
assert SEGMENT_DTYPE.itemsize == 48 and REPLICA_DTYPE.itemsize == 120

STATUS_MESSAGES = {
    -1: "replica exceeded the `steps` record capacity",
    -2: "nearest-hole search over zero holes",
    -3: "replay stream exhausted",
    -4: "allocation failed",
    -5: "lab row finished with zero steps",
    -6: "bad argument",
    -7: "CUDA error",
    -8: "replica does not fit the kernel's per-block capacity",
    -9: "kernel self-check failed",
}


def box_geometry(mc: Mapping[str, Any], phys: Mapping[str, Any],
                 e_ratio_start: Optional[float]) -> Tuple[float, int, int, int]:
    """(side, N_e, n_e0, n_h0) exactly as ``TLTrapSim.__init__`` + ``Box.seed`` compute them."""
    rho = mc["rho_prime"] * (3 / (4 * np.pi) * phys["alpha"] ** 3)      # tl_trap_lab.py:33
    side = (mc["holes"] / rho) ** (1 / 3)                                # tl_trap_lab.py:34
    e0 = int(mc["N_e"] * (e_ratio_start if e_ratio_start is not None else 0))   # :38
    total_h = int(mc["holes"] * mc["boundary_factor"] ** 3)              # engine.py:127
    return float(side), int(mc["N_e"]), e0, total_h


def legacy_box_geometry(mc: Mapping[str, Any], phys: Mapping[str, Any],
                        e_ratio_start: Optional[float]) -> Tuple[float, int, int, int]:
    """(side, N_e, n_e0, n_h0) as the LEGACY ``initialize_box_bg`` computes them
    (``src/est_params/functions.py:51-80``): the boundary shell adds ``int(density * (V_b - V))`` holes to
    ``int(holes)`` instead of ``int(holes * bf**3)``; same expressions, same operation order."""
    e0 = int(mc["N_e"] * (e_ratio_start if e_ratio_start is not None else 0))      # :57
    rho = mc["rho_prime"] * (3 / (4 * np.pi) * phys["alpha"] ** 3)                  # :82-86
    h = int(mc["holes"])                                                              # :60
    d = (h / rho) ** (1 / 3)                                                          # :61
    box_l, box_w, box_h = d, d, d
    b_factor = mc["boundary_factor"]
    box_l_b, box_w_b, box_h_b = box_l * b_factor, box_w * b_factor, box_h * b_factor
    box_volume = box_l * box_w * box_h
    box_volume_b = box_l_b * box_w_b * box_h_b
    holes_density = h / (box_l * box_w * box_h)
    holes_boundary_n = int(holes_density * (box_volume_b - box_volume))               # :74-76
    return float(d), int(mc["N_e"]), e0, h + holes_boundary_n


def fill_replica(rec: np.void, mc: Mapping[str, Any], phys_node: Mapping[str, Any],
                 e_ratio_start: Optional[float], protocol: int) -> None:
    phys = physics_record(phys_node)                 # TypeError on unknown keys, like Physics(**...)
    geometry = legacy_box_geometry if protocol == PROTO_TL_LEGACY else box_geometry
    side, N_e, e0, n_h0 = geometry(mc, phys, e_ratio_start)
    for k in ("alpha", "b", "s", "E_cb", "E_loc_1", "E_loc_2", "D0", "Retrap", "k_b"):
        rec[k] = float(phys[k])
    rec["side"] = side
    rec["boundary_factor"] = float(mc["boundary_factor"])
    rec["N_e"], rec["n_e0"], rec["n_h0"] = N_e, e0, n_h0
    rec["protocol"] = protocol
    rec["seg_begin"], rec["seg_count"] = 0, 1
    rec["obs_begin"], rec["obs_count"] = 0, 0


def simulate_tables(configs: Mapping[int, Mapping[str, Any]], sims: int):
    """Replica + segment tables for ``simulate()``: run-major, then sim (simulate.py:36,46).

    Returns ``(replicas[R], segments[n_runs])`` with ``R = n_runs * sims``.
    """
    n_runs = len(configs)
    reps = np.zeros(n_runs * sims, dtype=REPLICA_DTYPE)
    segs = np.zeros(n_runs, dtype=SEGMENT_DTYPE)
    for run_idx in range(n_runs):
        run_cfg = configs[run_idx]
        mc, phys = run_cfg["exp_type_fp"], run_cfg["physics_fp"]
        duration = float(mc.get("duration", 0.0))                  # simulate.py:40
        e_ratio_start = float(mc.get("e_ratio_start", 0.0))        # :41
        dose_rate = float(phys.get("dose_rate", 0.0))              # :42
        T_rate = float(mc.get("T_rate", 0.0))                      # :43
        max_dT = float(mc.get("max_dt", 0.5))                      # :44
        dt_cap = max_dT / T_rate if T_rate else 1e20               # :45
        T0 = mc.get("T_start", 0)                                  # :53 (getattr default 0)
        seg = segs[run_idx]
        seg["T_start"], seg["T_rate"], seg["duration"] = float(T0), T_rate, duration
        seg["dose_rate"], seg["dt_cap"], seg["A_opt"] = dose_rate, dt_cap, 0.0
        for j in range(sims):
            rec = reps[run_idx * sims + j]
            fill_replica(rec, mc, phys, e_ratio_start, PROTO_SIMULATE)
            rec["seg_begin"] = run_idx
    return reps, segs


# ---------------------------------------------------------------------------------------
# Lab protocols (Optimizer inner loop)
# ---------------------------------------------------------------------------------------
LAB_CSV = {                      # optimizer.py:73-78
    "iso": ("CLBR_IR50_ISO", PROTO_ISO_LAB),
    "tl_clbr": ("CLBR_IRSL50_0.25KperGy", PROTO_TL_LAB),
    "tl_fsm-13": ("FSM-13_IRSL50_0.25KperGy", PROTO_TL_LAB),
}


class LabTable:
    """One lab CSV turned into schedule rows (what TL_lab / ISO_lab loop over).

    TL  (tl_trap_lab.py:75-84): one row per CSV line; T_start = row.T_start, T_rate =
        (row.T_end - row.T_start) / row.Duration, e_ratio_start = row.e_ratio, target = row.Fill.
    ISO (tl_trap_lab.py:135-145,164-167): one row per exp_no; temperature, dose and e_ratio_start
        from the experiment's first line; observation times = its `time` column; the target of an
        observation is the e_ratio of the first line of the experiment with that time.
    """

    def __init__(self, csv_name: str, protocol: int, data_root: str):
        import os
        import pandas as pd

        self.csv_name, self.protocol = csv_name, protocol
        path = os.path.join(data_root, "data/processed", f"{csv_name}.csv")
        lab = pd.read_csv(path)
        self.frame = lab
        if protocol == PROTO_TL_LAB:
            n = len(lab)
            self.segments = np.zeros(n, dtype=SEGMENT_DTYPE)
            self.e_ratio_start = np.zeros(n)
            self.target = np.zeros(n)
            for k, row in lab.iterrows():
                seg = self.segments[k]
                seg["T_start"] = row.T_start
                seg["T_rate"] = (row.T_end - row.T_start) / row.Duration
                seg["duration"] = row.Duration
                seg["dt_cap"] = 1e20
                self.e_ratio_start[k] = row.e_ratio
                self.target[k] = row.Fill
            self.obs_begin = np.zeros(n + 1, dtype=np.int32)
            self.obs_time = np.zeros(0)
            self.exp_no = np.arange(n)
        else:
            exps = sorted(lab.exp_no.unique())
            n = len(exps)
            self.segments = np.zeros(n, dtype=SEGMENT_DTYPE)
            self.e_ratio_start = np.zeros(n)
            self.dose = np.zeros(n)
            obs_begin, obs_time, target = [0], [], []
            for k, exp_no in enumerate(exps):
                sub = lab[lab.exp_no == exp_no]
                seg = self.segments[k]
                seg["T_start"] = float(sub.temp.iloc[0])
                seg["dose_rate"] = float(sub.dose.iloc[0])
                seg["dt_cap"] = 1e20
                self.e_ratio_start[k] = float(sub.e_ratio.iloc[0])
                times = sub.time.to_numpy()
                for tt in times:
                    obs_time.append(float(tt))
                    target.append(float(sub[sub.time == tt].e_ratio.iloc[0]))
                obs_begin.append(len(obs_time))
            self.obs_begin = np.asarray(obs_begin, dtype=np.int32)
            self.obs_time = np.asarray(obs_time, dtype=np.float64)
            self.target = np.asarray(target, dtype=np.float64)
            self.exp_no = np.asarray(exps)
        self.n_rows = n

    def tables(self, run_cfg: Mapping[str, Any], legacy: bool = False):
        """Replica/segment tables for ONE parameter set: ``n_rows`` replicas.  ``legacy`` (TL only): the pre-refactor
        semantics of ``src/est_params/functions.py`` (protocol ``PROTO_TL_LEGACY``)."""
        mc, phys_node = run_cfg["exp_type_fp"], run_cfg["physics_fp"]
        reps = np.zeros(self.n_rows, dtype=REPLICA_DTYPE)
        segs = self.segments.copy()
        if legacy and self.protocol != PROTO_TL_LAB:
            raise ValueError("legacy semantics exist for the TL lab protocol only")
        if self.protocol == PROTO_TL_LAB:
            D = physics_record(phys_node)["D"]            # tl_trap_lab.py:83
            segs["dose_rate"] = float(D)                  # TypeError on None, like `D == 0` would not
        for k in range(self.n_rows):
            fill_replica(reps[k], mc, phys_node, float(self.e_ratio_start[k]), PROTO_TL_LEGACY if legacy else self.protocol)
            reps[k]["seg_begin"] = k
            reps[k]["obs_begin"] = self.obs_begin[k]
            reps[k]["obs_count"] = self.obs_begin[k + 1] - self.obs_begin[k]
        return reps, segs

    def mse(self, N_e: float, final_n_e: np.ndarray, obs_n_e: Optional[np.ndarray] = None):
        """(mean |err|, mse) exactly as tl_trap_lab.py:111-118 / :165-174 compute them."""
        if self.protocol == PROTO_TL_LAB:
            fills = [int(n) / N_e for n in final_n_e]
        else:
            fills = [int(n) / N_e for n in obs_n_e]
        err = [f - float(tg) for f, tg in zip(fills, self.target)]
        SE = [e ** 2 for e in err]
        ER = [abs(e) for e in err]
        return float(np.mean(ER)), float(np.mean(SE))
