"""Generate the golden vectors under ``tests/golden/`` from the UNMODIFIED reference.

Run in the build container only (needs ``/root/reference``):

    python oracle/ref_harness/gen_golden.py [--skip-kat0]

Every case seeds NumPy's legacy global stream (``np.random.seed``), calls the reference's own
``simulate(cfg)`` / ``optimizer.objective`` / ``optimizer.run_one_sim`` and stores what came
back: integer event / n_e traces, event times, the structural log ``(kind, electron index,
hole index)`` captured by wrapping ``Box.add_electron`` / ``Box.remove_pair``, and the number of
uniforms consumed.  The files are small (< 1 MB total) and are committed together with this
script; tests never need the reference tree.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, _HERE)
import refrun  # noqa: E402

GOLD = os.path.abspath(os.path.join(_HERE, "..", "..", "tests", "golden"))

SMALL = ["exp_type_fp.sims=1", "exp_type_fp.N_e=200", "exp_type_fp.holes=200", "exp_type_fp.steps=2000"]

SIM_CASES = {
    # SURVEY KAT-1
    "kat1": (SMALL + ["exp_type_fp.T_rate=[20]", "exp_type_fp.duration=[40]"], 0),
    # SURVEY KAT-2: isothermal, two tunnelling channels (the per-electron selector matters)
    "kat2": (SMALL + ["exp_type_fp.T_start=[250]", "exp_type_fp.T_rate=[0]", "exp_type_fp.duration=[1000]",
                      "physics_fp.E_loc_2=1.0", "physics_fp.Retrap=0.3"], 1),
    # 2 sweep points x 2 replicas on one continuing stream
    "sweep2x2": (["exp_type_fp.sims=2", "exp_type_fp.N_e=150", "exp_type_fp.holes=180",
                  "exp_type_fp.steps=1500", "exp_type_fp.T_rate=[5,20]", "exp_type_fp.duration=[60,20]"], 7),
    # partially filled traps, no boundary shell
    "partial": (["exp_type_fp.sims=1", "exp_type_fp.N_e=300", "exp_type_fp.holes=260", "exp_type_fp.steps=2000",
                 "exp_type_fp.e_ratio_start=0.3", "exp_type_fp.boundary_factor=[1.0]",
                 "exp_type_fp.T_rate=[10]", "exp_type_fp.duration=[70]"], 3),
    # empty box, no dose, no ramp: the spurious-fill corner of simulate.py:59-72
    "empty": (["exp_type_fp.sims=3", "exp_type_fp.N_e=50", "exp_type_fp.holes=60", "exp_type_fp.steps=100",
               "exp_type_fp.e_ratio_start=0", "exp_type_fp.T_rate=[0]", "exp_type_fp.duration=[100]"], 11),
    # finite E_cb (conduction-band channel active) with the simulate loop
    "labphys": (["physics_fp=lab_TL", "exp_type_fp.sims=2", "exp_type_fp.N_e=120", "exp_type_fp.holes=150",
                 "exp_type_fp.steps=3000", "exp_type_fp.T_rate=[2]", "exp_type_fp.duration=[250]",
                 "exp_type_fp.rho_prime=1e-5"], 5),
    # duration 0: `while t_cur <= 0` runs until the clock moves (simulate.py:51,91)
    "zero_duration": (SMALL + ["exp_type_fp.T_rate=[20]", "exp_type_fp.duration=[0]",
                               "exp_type_fp.T_start=[300]"], 4),
}

# first data row of the reference's results/lab_sims/result_tl_clbr.csv (legacy-code optimum)
def best_row():
    import pandas as pd
    df = pd.read_csv(os.path.join(refrun.REFERENCE_ROOT, "results", "lab_sims", "result_tl_clbr.csv"))
    return df.iloc[0].filter(like="param_").values.astype(float)


def sobol_candidates(n, seed):
    from scipy.stats import qmc
    m = refrun.load()
    bounds = np.array(m["optimizer"].DEFAULT_BOUNDS, dtype=float)
    pts = qmc.Sobol(d=10, seed=seed).random(n)
    return bounds[:, 0] + pts * (bounds[:, 1] - bounds[:, 0])


def trace_arrays(res):
    """Flatten simulate() outputs to per-replica traces in run-major, sim-minor order."""
    x_ax, lum, er = res["x_ax"], res["lum"], res["e_ratio"]
    steps, sims, runs = lum.shape
    configs = res["configs"]
    ev, ne, tt, used = [], [], [], []
    for run in range(runs):
        N_e = int(configs[run].exp_type_fp.N_e)
        for j in range(sims):
            t = x_ax[:, j, run]
            n = int(np.count_nonzero(t > 0))
            # records are contiguous from 0: every step moves the clock forward
            assert np.all(t[:n] > 0) and np.all(t[n:] == 0)
            used.append(n)
            ev.append(lum[:n, j, run].astype(np.int8))
            ne.append(np.rint(er[:n, j, run] * N_e).astype(np.int32))
            tt.append(t[:n].copy())
    return used, ev, ne, tt


def save_sim_case(name, overrides, seed):
    t0 = time.time()
    res = refrun.run_simulate(overrides, seed)
    used, ev, ne, tt = trace_arrays(res)
    log = np.asarray(res["rec"].events, dtype=np.int32).reshape(-1, 3)
    np.savez_compressed(
        os.path.join(GOLD, f"sim_{name}.npz"),
        steps_used=np.asarray(used, np.int32),
        event=np.concatenate(ev) if ev else np.zeros(0, np.int8),
        n_e=np.concatenate(ne) if ne else np.zeros(0, np.int32),
        t=np.concatenate(tt) if tt else np.zeros(0),
        log=log,
    )
    meta = dict(kind="simulate", overrides=list(overrides), seed=seed,
                n_uniforms=int(res["rec"].n_uniforms), steps_used=[int(u) for u in used],
                events=[int(e.sum()) for e in ev], wall_s=round(time.time() - t0, 2))
    print(name, meta["steps_used"], meta["events"], meta["wall_s"], "s")
    return meta


def save_lab_case(name, exp, seed, p=None):
    t0 = time.time()
    if p is None:
        res = refrun.run_lab(exp, seed)
    else:
        res = refrun.run_objective(p, exp, seed)
    log = np.asarray(res["rec"].events, dtype=np.int32).reshape(-1, 3)
    np.savez_compressed(os.path.join(GOLD, f"lab_{name}.npz"), log=log,
                        p=np.asarray(p if p is not None else [], dtype=np.float64))
    meta = dict(kind="lab", exp=exp, seed=seed, value=res["value"],
                p=[float(v) for v in p] if p is not None else None,
                n_uniforms=int(res["rec"].n_uniforms), n_events=int(log.shape[0]),
                printed=res["printed"].strip(), wall_s=round(time.time() - t0, 2))
    print(name, exp, repr(res["value"]), log.shape[0], meta["wall_s"], "s")
    return meta


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-kat0", action="store_true", help="skip the 3-minute default-config case")
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    man_path = os.path.join(GOLD, "manifest.json")
    manifest = json.load(open(man_path)) if os.path.isfile(man_path) else {}

    def want(n):
        return args.only is None or args.only == n

    for name, (ov, seed) in SIM_CASES.items():
        if want(name):
            manifest[f"sim_{name}"] = save_sim_case(name, ov, seed)

    lab_cases = [("default_tl_clbr", "tl_clbr", 2, None), ("default_tl_fsm13", "tl_fsm-13", 2, None),
                 ("default_iso", "iso", 2, None)]
    bp = best_row()
    lab_cases += [("kat3_tl_clbr", "tl_clbr", 2, bp), ("kat3_tl_fsm13", "tl_fsm-13", 2, bp),
                  ("kat3_iso", "iso", 2, bp)]
    cand = sobol_candidates(4, seed=4)
    for k in range(4):
        lab_cases.append((f"sobol{k}_tl_clbr", "tl_clbr", 100 + k, cand[k]))
    for k in range(2):
        lab_cases.append((f"sobol{k}_iso", "iso", 200 + k, cand[k]))
    for name, exp, seed, p in lab_cases:
        if want(name):
            manifest[f"lab_{name}"] = save_lab_case(name, exp, seed, p)

    # legacy (pre-refactor) TL code, src/est_params/functions.py: what produced results/lab_sims/result_tl_clbr.csv
    legacy_cases = [("default", None, 3), ("default_b", None, 4), ("best_row", bp, 3), ("best_row_b", bp, 5),
                    ("sobol0", cand[0], 300), ("sobol1", cand[1], 301), ("sobol2", cand[2], 302)]
    for name, p, seed in legacy_cases:
        if want("legacy_" + name):
            t0 = time.time()
            try:
                res = refrun.run_legacy_tl(p, seed)
            except Exception as exc:  # noqa: BLE001  (a candidate the legacy code itself cannot finish)
                print("legacy", name, "failed in the reference:", type(exc).__name__, exc)
                continue
            np.savez_compressed(os.path.join(GOLD, f"legacy_{name}.npz"), rows=res["rows"],
                                p=np.asarray(p if p is not None else [], dtype=np.float64))
            manifest[f"legacy_{name}"] = dict(kind="legacy_tl", seed=seed, value=res["value"],
                                              p=[float(v) for v in p] if p is not None else None,
                                              wall_s=round(time.time() - t0, 2))
            print("legacy", name, repr(res["value"]), res["rows"][:, 2:].sum(0), manifest[f"legacy_{name}"]["wall_s"], "s")

    # SURVEY KAT-0: the shipped default config (TL12 + basicTL12), ~3 minutes on one core
    if not args.skip_kat0 and want("kat0"):
        manifest["sim_kat0"] = save_sim_case("kat0", [], 12345)

    with open(man_path, "w") as fh:
        json.dump(manifest, fh, indent=1, sort_keys=True)
    print("wrote", man_path)


if __name__ == "__main__":
    main()
