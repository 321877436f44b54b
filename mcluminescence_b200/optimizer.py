"""Optimizer call contract: ``objective(p, cfg, exp)`` and a batched form for SciPy DE.

Mirrors the reference's ``src/class/optimizer.py``: ``DEFAULT_BOUNDS`` (:32-43),
``cfg_with_params`` (:49-65), ``run_one_sim`` (:68-80), ``objective`` (:82-84) and the
train / replay dispatcher (:89-137).  The simulation itself runs on the GPU through the C ABI.
"""
from __future__ import annotations

from typing import Any, List, Mapping, Tuple

import numpy as np

DEFAULT_BOUNDS: List[Tuple[float, float]] = [
    (1e-8, 1e-3),     # rho_prime
    (1.9, 2.4),       # E_cb
    (1.2, 1.7),       # E_loc_1
    (1.0, 1.5),       # E_loc_2
    (1e2, 4e2),       # D0
    (1e12, 1e14),     # s
    (1e10, 1e13),     # b
    (1e9, 5e10),      # alpha
    (1e2, 7.5e2),     # holes
    (0.0, 1.0),       # retrap
]

PARAM_ORDER = ("rho_prime", "E_cb", "E_loc_1", "E_loc_2", "D0", "s", "b", "alpha", "holes", "retrap")


def cfg_with_params(base: Mapping[str, Any], params: np.ndarray):
    """Copy of *base* with the ten fitted fields replaced (reference optimizer.py:49-65).

    Like the reference the copy is shallow: the nested ``exp_type_fp`` / ``physics_fp`` nodes
    are shared with *base*; all ten fields are overwritten on every call so this is harmless.
    """
    cfg = base.copy()
    (rho_prime, E_cb, E_loc_1, E_loc_2, D0, s, b, alpha, holes, retrap) = params
    cfg.exp_type_fp.rho_prime = float(rho_prime)
    cfg.exp_type_fp.holes = float(holes)
    cfg.physics_fp.E_cb = float(E_cb)
    cfg.physics_fp.E_loc_1 = float(E_loc_1)
    cfg.physics_fp.E_loc_2 = float(E_loc_2)
    cfg.physics_fp.D0 = float(D0)
    cfg.physics_fp.s = float(s)
    cfg.physics_fp.b = float(b)
    cfg.physics_fp.alpha = float(alpha)
    cfg.physics_fp.Retrap = float(retrap)
    return cfg


def run_one_sim(cfg: Mapping[str, Any], exp: str, PLOT: bool = False, **kw) -> float:
    """MSE for one parameter set and one experiment label (reference optimizer.py:68-80)."""
    from .config import initialize_runs
    from .tl_trap_lab import TLTrapSim
    runs = initialize_runs(cfg)
    for k in ("rng", "seed"):
        if k in cfg and k not in runs[0]:
            runs[0][k] = cfg[k]
    sim = TLTrapSim(runs[0], **kw)
    if exp == "iso":
        return sim.ISO_lab("CLBR_IR50_ISO", plot=PLOT)
    elif exp == "tl_clbr":
        return sim.TL_lab("CLBR_IRSL50_0.25KperGy", plot=PLOT)
    elif exp == "tl_fsm-13":
        return sim.TL_lab("FSM-13_IRSL50_0.25KperGy", plot=PLOT)
    else:
        raise ValueError(f"Unknown exp '{exp}'")


def objective(p: np.ndarray, cfg: Mapping[str, Any], exp: str, **kw) -> float:
    """Reference optimizer.py:82-84."""
    return run_one_sim(cfg_with_params(cfg, p), exp, **kw)
