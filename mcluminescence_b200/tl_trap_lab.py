"""``TLTrapSim`` -- drop-in for the reference's ``src/class/tl_trap_lab.py`` lab protocols.

``TLTrapSim(cfg, e_ratio_start=None)`` validates the physics node and derives the box geometry
like the reference constructor (``tl_trap_lab.py:27-43``); ``TL_lab(csv)`` / ``ISO_lab(csv)`` run
every lab row / isothermal experiment of the CSV as one GPU replica each and return the same
mean squared error (``tl_trap_lab.py:65-123,125-179``), printing the same summary line.

The per-step Python helpers of the reference (``_update_lifetimes``, ``_filling_time``,
``box.add_electron`` ...) are NOT a device boundary here: a step is a few hundred nanoseconds on
the GPU, so the whole row loop lives in the kernel.
"""
from __future__ import annotations

import os
from types import SimpleNamespace
from typing import Any, Mapping, Optional

import numpy as np

from . import engine
from .config import DATA_DIR, physics_record
from .replicas import (LabTable, MODE_PHILOX, PROTO_ISO_LAB, PROTO_TL_LAB, box_geometry)

#: root that holds ``data/processed/<csv>.csv`` (the reference uses its repo root)
PROJECT_ROOT = os.path.dirname(DATA_DIR)

_lab_cache = {}


def lab_table(csv_name: str, protocol: int, root: Optional[str] = None) -> LabTable:
    key = (root or PROJECT_ROOT, csv_name, protocol)
    if key not in _lab_cache:
        _lab_cache[key] = LabTable(csv_name, protocol, key[0])
    return _lab_cache[key]


class TLTrapSim:
    def __init__(self, cfg: Mapping[str, Any], *, e_ratio_start: Optional[float] = None,
                 rng: Optional[str] = None, seed: Optional[int] = None, candidate_id: int = 0, device=None,
                 legacy: Optional[bool] = None):
        self.cfg = cfg
        self.mc = cfg["exp_type_fp"]
        self.phys = SimpleNamespace(**physics_record(cfg["physics_fp"]))   # TypeError like Physics(**...)
        side, N_e, e0, n_h0 = box_geometry(self.mc, vars(self.phys), e_ratio_start)
        self.box = SimpleNamespace(L=side, W=side, H=side, boundary_factor=self.mc["boundary_factor"],
                                   n_e0=e0, n_h0=n_h0)
        self.rng = rng if rng is not None else str(cfg.get("rng", "philox"))
        self.seed = seed if seed is not None else cfg.get("seed", None)
        self.device = device
        self.candidate_id = int(candidate_id)     # global id of this parameter set (Philox stream key)
        # legacy=True: TL_lab follows the PRE-REFACTOR code (reference src/est_params/functions.py:270-360, what produced
        # results/lab_sims/result_*.csv) instead of src/class/tl_trap_lab.py -- see MCL_PROTO_TL_LEGACY in mcl_b200.h
        self.legacy = bool(legacy if legacy is not None else cfg.get("legacy", False))
        if self.legacy and self.rng == "replay":
            raise ValueError("the legacy semantics run in native (philox) mode only")
        self.last_esteps = 0
        if self.rng == "replay":
            # the reference constructor seeds a Box: 3*(e0 + n_h0) uniforms leave the global stream
            if self.seed is not None:
                engine.seed_replay(int(self.seed))
            engine.global_replay().advance(3 * (e0 + n_h0))

    # ------------------------------------------------------------------
    def _run(self, lt: LabTable):
        reps, segs = lt.tables(self.cfg, legacy=self.legacy)
        steps = int(self.mc["steps"])
        if self.rng == "replay":
            res = engine.run_replay_chained(reps, segs, steps, engine.global_replay(),
                                            obs_time=lt.obs_time, device=self.device)
            status = res["status"]
            final_n_e, obs_n_e, esteps = res["final_n_e"], res["obs_n_e"], res["esteps"]
        else:
            seed = self.seed if self.seed is not None else int.from_bytes(os.urandom(8), "little")
            out = engine.run_replicas(reps, segs, steps, mode=MODE_PHILOX, seed=int(seed),
                                      replica_id0=self.candidate_id * lt.n_rows,
                                      obs_time=lt.obs_time, trace=False, device=self.device)
            status = out.status
            final_n_e, obs_n_e, esteps = out.final_n_e, out.obs_n_e, out.esteps
        bad = np.nonzero(status)[0]
        if bad.size:
            code = int(status[bad[0]])
            if code in (-1, -5):
                raise IndexError("list index out of range" if code == -5 else
                                 f"lab row {int(bad[0])} exceeded steps={steps}")
            raise RuntimeError(f"lab row {int(bad[0])} failed with status {code}")
        self.last_esteps = int(np.sum(esteps))
        return final_n_e, obs_n_e

    def TL_lab(self, csv_name: str, *, plot: bool = False) -> float:
        if plot:
            raise NotImplementedError("plotting is outside the kinetics hot path")
        lt = lab_table(csv_name, PROTO_TL_LAB)
        final_n_e, _ = self._run(lt)
        avg_er, mse = lt.mse(self.mc["N_e"], final_n_e)
        self._printer(avg_er, mse)
        return mse

    def ISO_lab(self, csv_name: str, *, plot: bool = False) -> float:
        if plot:
            raise NotImplementedError("plotting is outside the kinetics hot path")
        lt = lab_table(csv_name, PROTO_ISO_LAB)
        _, obs_n_e = self._run(lt)
        avg_er, mse = lt.mse(self.mc["N_e"], None, obs_n_e)
        self._printer(avg_er, mse)
        return mse

    def _printer(self, avgER: float, MSE: float) -> None:       # tl_trap_lab.py:185-190
        p = self.cfg["physics_fp"]
        mc = self.mc
        print(f"absError={avgER:.3e}, MSE={MSE:.3e}  [rho'={mc['rho_prime']}, E_cb={p['E_cb']}, D0={p['D0']}, "
              f"E_loc1={p['E_loc_1']}, E_loc2={p['E_loc_2']}, s={p['s']}, b={p['b']}, alpha={p['alpha']}, "
              f"holes={mc['holes']}, P_retrap={p['Retrap']}]")
