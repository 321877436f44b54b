"""Shared helpers for the parity tests: golden-vector access and table building."""
from __future__ import annotations

import json
import os

import numpy as np

from mcluminescence_b200.config import DATA_DIR, compose, initialize_runs
from mcluminescence_b200.replicas import LAB_CSV, LabTable, simulate_tables

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DATA_ROOT = os.path.dirname(DATA_DIR)

LAB_OVERRIDES = ["exp_type_fp=TLlab", "physics_fp=lab_TL"]


class Golden:
    def __init__(self):
        with open(os.path.join(GOLD, "manifest.json")) as fh:
            self.manifest = json.load(fh)

    def names(self, kind):
        return sorted(k for k, v in self.manifest.items() if v["kind"] == kind)

    def meta(self, name):
        return self.manifest[name]

    def arrays(self, name):
        return np.load(os.path.join(GOLD, f"{name}.npz"))


def sim_tables_for(meta):
    """(replicas, segments, max_steps, sims, n_runs) for a golden simulate case, built through the
    PRODUCT's own conf/ copy -- so these tests also pin the shipped YAML surface."""
    cfg = compose(overrides=meta["overrides"])
    runs = initialize_runs(cfg)
    mc0 = runs[0]["exp_type_fp"]
    sims, steps = int(mc0["sims"]), int(mc0["steps"])
    reps, segs = simulate_tables(runs, sims)
    return reps, segs, steps, sims, len(runs)


def split_traces(arrs):
    """Per-replica (event, n_e, t) from the concatenated golden arrays."""
    used = arrs["steps_used"]
    off = np.concatenate([[0], np.cumsum(used)])
    return [(arrs["event"][off[i]:off[i + 1]].astype(np.int32),
             arrs["n_e"][off[i]:off[i + 1]].astype(np.int32),
             arrs["t"][off[i]:off[i + 1]]) for i in range(len(used))]


def lab_setup(meta):
    """(LabTable, replicas, segments, run_cfg) for a golden lab case."""
    from mcluminescence_b200.optimizer import cfg_with_params
    cfg = compose(overrides=LAB_OVERRIDES)
    if meta["p"] is not None:
        cfg = cfg_with_params(cfg, np.asarray(meta["p"], dtype=float))
    run = initialize_runs(cfg)[0]
    csv, proto = LAB_CSV[meta["exp"]]
    lt = LabTable(csv, proto, DATA_ROOT)
    reps, segs = lt.tables(run)
    return lt, reps, segs, run


def log_from_result(res, r_list):
    """Structural log rows (kind, e_idx, h_idx) with kind 1 fill / 2 recombination, replica order."""
    rows = []
    for r in r_list:
        n = int(res.steps_used[r])
        k = res.kind[r, :n]
        m = k > 0
        rows.append(np.stack([k[m], res.e_idx[r, :n][m], res.h_idx[r, :n][m]], axis=1))
    return rows
