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


def legacy_setup(meta):
    """(LabTable, replicas, segments, run_cfg) for a golden case of the LEGACY TL code (protocol PROTO_TL_LEGACY)."""
    from mcluminescence_b200.optimizer import cfg_with_params
    cfg = compose(overrides=LAB_OVERRIDES)
    if meta["p"] is not None:
        cfg = cfg_with_params(cfg, np.asarray(meta["p"], dtype=float))
    run = initialize_runs(cfg)[0]
    csv, proto = LAB_CSV["tl_clbr"]
    lt = LabTable(csv, proto, DATA_ROOT)
    reps, segs = lt.tables(run, legacy=True)
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


# ---------------------------------------------------------------------------------------
# Per-leg statistics of a (event, n_e, t) trace on the kernel's histogram axes, restated in NumPy
# ---------------------------------------------------------------------------------------
def split_legs(t, durs):
    """[(i0, i1, t_off)] per leg of ONE replica trace.  A leg ends with the step whose local time reaches its
    duration (simulate.py:91-92); the next leg restarts its clock (records carry t_off + t_local)."""
    legs, i0, t_off = [], 0, 0.0
    for d in durs:
        loc = t[i0:] - t_off
        k = int(np.searchsorted(loc, d, side="left"))         # first step with t_local >= duration
        i1 = min(i0 + k + 1, len(t))
        legs.append((i0, i1, t_off))
        if i1 > i0:
            t_off = float(t[i1 - 1])
        i0 = i1
    return legs


def axis_edges_in_time(hist, seg):
    """The n_bins + 1 bin edges of the kernel's histogram axis, as LOCAL TIMES of the leg `seg`."""
    if hist.axis == 1:                                        # MCL_AXIS_TIME_LOG: same recurrence as the kernel
        ratio = 10.0 ** ((np.log10(hist.hi) - np.log10(hist.lo)) / hist.n_bins)
        e = [hist.lo]
        for _ in range(hist.n_bins):
            e.append(e[-1] * ratio)
        return np.array(e)
    v = hist.lo + np.arange(hist.n_bins + 1) * ((hist.hi - hist.lo) / hist.n_bins)
    if hist.axis == 2:                                        # MCL_AXIS_TEMP on a heating leg
        return (v - float(seg["T_start"])) / float(seg["T_rate"])
    return v


def leg_histograms(event, n_e, t, used, n_e0, segments, hist):
    """Per replica and leg: (events per bin, occupancy at each bin's left edge [NaN = edge never crossed], n_e at the
    end of the leg, events in the leg, local event times) following the kernel's histogram definition."""
    R, n_seg = event.shape[0], len(segments)
    ev = np.zeros((R, n_seg, hist.n_bins))
    occ = np.full((R, n_seg, hist.n_bins), np.nan)
    n_end = np.zeros((R, n_seg))
    n_ev = np.zeros((R, n_seg))
    times = [[] for _ in range(n_seg)]
    for r in range(R):
        tt, ee, nn = t[r, :used[r]], event[r, :used[r]], n_e[r, :used[r]]
        n_start = n_e0
        for sg, (i0, i1, t_off) in enumerate(split_legs(tt, segments["duration"])):
            edges = axis_edges_in_time(hist, segments[sg])
            loc = tt[i0:i1] - t_off
            hit = ee[i0:i1] > 0
            k = np.searchsorted(edges, loc[hit], side="right") - 1
            k = k[(k >= 0) & (k < hist.n_bins)]
            np.add.at(ev[r, sg], k, 1.0)
            idx = np.searchsorted(loc, edges[:-1], side="left")       # the step that crosses each left edge
            n_before = np.concatenate([[n_start], nn[i0:i1]])[np.minimum(idx, len(loc))]
            occ[r, sg] = np.where(idx < len(loc), n_before, np.nan)
            n_end[r, sg] = nn[i1 - 1] if i1 > i0 else n_start
            n_ev[r, sg] = hit.sum()
            times[sg].append(loc[hit])
            n_start = int(n_end[r, sg])
    return ev, occ, n_end, n_ev, [np.concatenate(x) if x else np.zeros(0) for x in times]
