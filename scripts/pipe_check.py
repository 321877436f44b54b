"""Pipelined vs in-order specialised step loop (MCL_PHILOX_PIPE=1 / 0): identical fingerprints, then throughput A/B.

    python scripts/pipe_check.py [replicas for the timing, default 2960]
"""
import os, sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
from mcluminescence_b200 import engine, workloads


def run(wl, pipe, seed=43, nt=None):
    os.environ["MCL_PHILOX_PIPE"] = "1" if pipe else "0"
    if nt:
        os.environ["MCL_PHILOX_NT"] = str(nt)
    try:
        return engine.run_replicas(wl["replicas"], wl["segments"], wl["max_steps"], seed=seed, hist=wl.get("hist"), trace=False, sync=True)
    finally:
        os.environ.pop("MCL_PHILOX_NT", None)


def same(a, b, tag):
    bad = [k for k in ("status", "steps_used", "final_n_e", "esteps", "hist_events", "hist_occ")
           if getattr(a, k, None) is not None and not np.array_equal(np.asarray(getattr(a, k)), np.asarray(getattr(b, k)))]
    ndiff = int((np.asarray(a.steps_used) != np.asarray(b.steps_used)).sum())
    print(f"{tag}: {'IDENTICAL' if not bad else 'DIFFERENT ' + str(bad)}  status {np.unique(np.asarray(a.status)).tolist()} replicas differing {ndiff}/{len(a.steps_used)}", flush=True)
    return not bad


ok = True
jobs = [("c2 x12 nt256", workloads.c2(n_replicas=12), 256), ("c2 x6 nt512", workloads.c2(n_replicas=6), None),
        ("c2 two-channel 5000e", workloads.c2(n_replicas=6, n_e=5000, physics_overrides=["physics_fp.E_loc_2=1.0", "physics_fp.Retrap=0.3"]), 256),
        ("c2 2000e nt128", workloads.c2(n_replicas=24, n_e=2000), 128), ("c2 x300", workloads.c2(n_replicas=300), None),
        ("c5 x300 nt64", workloads.c5(n_replicas=300), 64), ("c5 x300 nt128", workloads.c5(n_replicas=300), 128), ("c5 x100 nt256", workloads.c5(n_replicas=100), 256)]
if os.environ.get("ONLY256"):          # libraries built with -DMCL_ONLY_C2 hold the 256-thread kernel only
    jobs = [j for j in jobs if j[2] == 256 or j[0] == "c2 x300"]
for tag, wl, nt in jobs:
    a = run(wl, True, nt=nt); b = run(wl, False, nt=nt)
    ok &= same(a, b, tag)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2960
wl = workloads.c2(n_replicas=n)
nt_t = int(os.environ["TIME_NT"]) if os.environ.get("TIME_NT") else None
for pipe in (1, 0, 1, 0):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    o = run(wl, pipe, seed=7, nt=nt_t)
    dt = time.perf_counter() - t0
    print(f"pipe={pipe}: {int(np.asarray(o.esteps).sum()) / dt / 1e9:.1f} G electron-steps/s wall ({dt * 1e3:.0f} ms)", flush=True)
sys.exit(0 if ok else 1)
