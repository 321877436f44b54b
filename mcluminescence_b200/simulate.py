"""``simulate(cfg)`` -- drop-in for the reference's ``src/class/simulate.py``.

Same call contract: takes the composed config, returns ``(x_ax, Lum, electron_ratio, configs)``
with arrays of shape ``[steps, sims, runs]`` (float64, zero padded past each replica's last
step) and writes ``results/simulations/exp_{tag}.csv`` with columns
``run,sim,step,lum,electron_ratio`` (``simulate.py:95-112``).  The per-replica Monte-Carlo loop
(``simulate.py:46-92``) runs on the GPU, all sweep points x replicas in one launch.

Extra top-level flags (read with ``cfg.get`` like the reference's ``tag``):
  ``rng``   ``philox`` (default) or ``replay`` (bit-exact against the reference given ``seed``)
  ``seed``  integer; Philox default is drawn from the OS (the reference is unseeded)
"""
from __future__ import annotations

import os
import sys
from typing import Any, Mapping, Optional

import numpy as np

from . import engine
from .config import compose, initialize_runs
from .replicas import MODE_PHILOX, MODE_REPLAY, simulate_tables

#: where ``results/simulations/exp_{tag}.csv`` goes (the reference uses its repo root)
PROJECT_ROOT = os.getcwd()


def simulate(cfg: Mapping[str, Any], *, rng: Optional[str] = None, seed: Optional[int] = None,
             write_csv: bool = True, device=None):
    configs = initialize_runs(cfg)                       # simulate.py:22
    n_runs = len(configs)
    mc0 = configs[0]["exp_type_fp"]
    steps = int(mc0["steps"])                            # :26
    sims = int(mc0["sims"])                              # :27
    tag = cfg.get("tag", "")                             # :33
    rng = rng if rng is not None else str(cfg.get("rng", "philox"))
    if seed is None:
        seed = cfg.get("seed", None)

    reps, segs = simulate_tables(configs, sims)
    if rng == "replay":
        stream = engine.seed_replay(int(seed)) if seed is not None else engine.global_replay()
        res = engine.run_replay_chained(reps, segs, steps, stream, device=device)
        status, used = res["status"], res["steps_used"]
        ev, ne, tt = res["event"], res["n_e"], res["t"]
        bad = np.nonzero(status)[0]
        if bad.size:
            if int(status[bad[0]]) == -1:
                raise IndexError(f"index {steps} is out of bounds for axis 0 with size {steps}")
            raise RuntimeError(f"replay replica {int(bad[0])} failed with status {int(status[bad[0]])}")
    elif rng == "philox":
        if seed is None:
            seed = int.from_bytes(os.urandom(8), "little")
        out = engine.run_replicas(reps, segs, steps, mode=MODE_PHILOX, seed=int(seed), device=device)
        out.raise_on_error()
        ev, ne, tt, used = out.event, out.n_e, out.t, out.steps_used
    else:
        raise ValueError(f"rng must be 'philox' or 'replay', got {rng!r}")

    # [R, steps] -> [steps, sims, runs]; replica index = run * sims + sim
    N_e = np.array([configs[r]["exp_type_fp"]["N_e"] for r in range(n_runs)], dtype=np.float64)
    mask = np.arange(steps)[None, :] < used[:, None]
    x_rs = np.where(mask, tt, 0.0).reshape(n_runs, sims, steps)
    lum_rs = np.where(mask, ev, 0).astype(np.float64).reshape(n_runs, sims, steps)
    er_rs = (np.where(mask, ne, 0).reshape(n_runs, sims, steps) / N_e[:, None, None])
    x_ax = np.ascontiguousarray(x_rs.transpose(2, 1, 0))
    Lum = np.ascontiguousarray(lum_rs.transpose(2, 1, 0))
    electron_ratio = np.ascontiguousarray(er_rs.transpose(2, 1, 0))

    if write_csv:
        import pandas as pd
        df = pd.DataFrame({
            "run": np.repeat(np.arange(n_runs), sims * steps),
            "sim": np.tile(np.repeat(np.arange(sims), steps), n_runs),
            "step": np.tile(np.arange(steps), n_runs * sims),
            "lum": lum_rs.ravel(),
            "electron_ratio": er_rs.ravel(),
        })
        out_dir = os.path.join(PROJECT_ROOT, "results", "simulations")
        os.makedirs(out_dir, exist_ok=True)
        df.to_csv(os.path.join(out_dir, f"exp_{tag}.csv"), index=False)
    return x_ax, Lum, electron_ratio, configs


def main(argv=None):
    """``python -m mcluminescence_b200.simulate [overrides...]`` (Hydra-style overrides)."""
    cfg = compose("config_fp", list(sys.argv[1:] if argv is None else argv))
    return simulate(cfg)


if __name__ == "__main__":
    main()
