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


def objective_batched(P: np.ndarray, cfg: Mapping[str, Any], exp: str, *, seed=None,
                      candidate_id0: int = 0, return_esteps: bool = False, legacy: bool = False):
    """``objective`` for a whole population: ``P[10, S]`` -> ``mse[S]``, one GPU launch.

    This is the seam SciPy offers with ``differential_evolution(..., vectorized=True,
    updating="deferred")``: the solver hands ``x`` of shape ``(N, S)`` and expects ``(S,)``.
    Candidate ``c`` is evaluated exactly like ``objective(P[:, c], cfg, exp, seed=seed,
    candidate_id=candidate_id0 + c)`` (same Philox streams), replacing S passes through
    reference ``optimizer.py:82-84``.  Candidates the reference would crash on get ``inf``.
    """
    import ctypes as C
    import os

    from . import _native
    from .config import initialize_runs, physics_record
    from .engine import _torch
    from .replicas import LAB_CSV, PROTO_ISO_LAB
    from .tl_trap_lab import lab_table

    if exp not in LAB_CSV:
        raise ValueError(f"Unknown exp '{exp}'")
    P = np.ascontiguousarray(P, dtype=np.float64)
    if P.ndim != 2 or P.shape[0] != 10:
        raise ValueError("P must have shape (10, S)")
    S = int(P.shape[1])
    torch = _torch()
    L = _native.load()
    run = initialize_runs(cfg)[0]
    mc, phys = run["exp_type_fp"], physics_record(run["physics_fp"])
    csv, proto = LAB_CSV[exp]
    lt = lab_table(csv, proto)
    rows = np.ascontiguousarray(lt.segments)
    ers = np.ascontiguousarray(lt.e_ratio_start, dtype=np.float64)
    tgt = np.ascontiguousarray(lt.target, dtype=np.float64)
    obs_b = np.ascontiguousarray(lt.obs_begin, dtype=np.int32)
    obs_t = np.ascontiguousarray(lt.obs_time, dtype=np.float64)
    lab = _native.Lab()
    lab.protocol, lab.n_rows = proto, lt.n_rows
    lab.rows, lab.e_ratio_start, lab.target = rows.ctypes.data, ers.ctypes.data, tgt.ctypes.data
    if proto == PROTO_ISO_LAB:
        lab.obs_begin, lab.obs_time = obs_b.ctypes.data, obs_t.ctypes.data
    lab.N_e, lab.boundary_factor = float(mc["N_e"]), float(mc["boundary_factor"])
    lab.D, lab.k_b = float(phys["D"] if phys["D"] is not None else 0.0), float(phys["k_b"])
    lab.max_steps = int(mc["steps"])
    lab.flags = 1 if (legacy or cfg.get("legacy", False)) else 0       # MCL_LAB_LEGACY: pre-refactor est_params semantics (TL only)
    if seed is None:
        seed = cfg.get("seed", None)
    if seed is None:
        seed = int.from_bytes(os.urandom(8), "little")
    mse = np.zeros(S, dtype=np.float64)
    est = C.c_int64(0)
    stream = torch.cuda.current_stream().cuda_stream
    rc = L.mcl_objective(P.ctypes.data, S, C.byref(lab), int(seed) & (2 ** 64 - 1), int(candidate_id0),
                         mse.ctypes.data, C.byref(est), stream)
    _native.check(rc, "mcl_objective")
    return (mse, int(est.value)) if return_esteps else mse


#: root for ``results/lab_sims/result_{exp}.csv`` (the reference uses its repo root)
PROJECT_ROOT = None


def main(cfg: Mapping[str, Any]) -> None:
    """Train / replay dispatcher with the reference's flags and CSV format (optimizer.py:89-137).

    ``task=train``: SciPy differential evolution with the reference's settings, except that the
    population is evaluated through ``objective_batched`` (``vectorized=True``,
    ``updating="deferred"``) instead of a process pool (``workers``): one GPU launch per generation.
    ``task=replay``: re-simulate the best stored row.
    """
    import os
    from pathlib import Path

    import pandas as pd
    from scipy.optimize import differential_evolution

    task = cfg.get("task", "train")
    exp = cfg.get("exp", "tl_clbr")
    gens = int(cfg.get("gens", 7))
    pop = int(cfg.get("pop", 15))
    root = PROJECT_ROOT or os.getcwd()
    csv_path = Path(root, f"results/lab_sims/result_{exp}.csv")
    csv_path.parent.mkdir(parents=True, exist_ok=True)

    if task == "train":
        calls = {"n": 0}

        def batched(X):
            X = np.asarray(X, dtype=float)
            single = X.ndim == 1                    # the final L-BFGS-B polish evaluates one point at a time
            P = X[:, None] if single else X
            out = objective_batched(P, cfg, exp, candidate_id0=calls["n"])
            calls["n"] += P.shape[1]
            return float(out[0]) if single else out

        result = differential_evolution(
            batched, DEFAULT_BOUNDS, strategy="randtobest1bin", init="sobol", mutation=0.5,
            recombination=0.3, maxiter=gens, popsize=pop, tol=1e-7, disp=True, polish=True,
            vectorized=True, updating="deferred")
        cols = [f"param_{i}" for i in range(len(result.x))] + ["mse"]
        df = pd.DataFrame([list(result.x) + [float(result.fun)]], columns=cols)
        df.to_csv(csv_path, mode="a", index=False, header=not csv_path.exists())
        print(f"Saved run to {csv_path}")
        run_one_sim(cfg_with_params(cfg, result.x), exp)
    elif task == "replay":
        if not csv_path.exists():
            raise FileNotFoundError(csv_path)
        df = pd.read_csv(csv_path)
        best_row = df.loc[df.mse.idxmin()]
        best_p = best_row.filter(like="param_").values.astype(float)
        run_one_sim(cfg_with_params(cfg, best_p), exp)
        print("Re-simulated best parameters for exp:", exp)
    else:
        raise ValueError("task must be 'train' or 'replay'")


if __name__ == "__main__":
    import sys
    from .config import compose
    main(compose("config_fp", ["exp_type_fp=TLlab", "physics_fp=lab_TL"] + sys.argv[1:]))
