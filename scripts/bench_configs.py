"""Throughput of the other BASELINE configs (C1, C3, C4, C5) on one GPU, for the docs.  Not the bench contract."""
import json, sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from mcluminescence_b200 import engine, ensemble, optimizer, workloads
from mcluminescence_b200.config import compose

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9; out = None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best, out

res = {}
for name, wl in (("c1", workloads.c1()), ("c3", workloads.c3(replicas_per_dose=256)), ("c5", workloads.c5(n_replicas=6250))):
    def go():
        fin, T = ensemble.run_ensemble(wl, seed=7)
        return T
    t, T = timed(go)
    es = int(T["counters"][0].item()); errs = int(T["counters"][2].item())
    res[name] = dict(workload=wl["name"], seconds=t, esteps=es, esteps_per_s=es / t, errors=errs)
cfg = compose(overrides=["exp_type_fp=TLlab", "physics_fp=lab_TL"])
P = workloads.c4_candidates(4096, seed=4)
for exp in ("tl_clbr", "iso"):
    t0 = time.perf_counter(); mse, es = optimizer.objective_batched(P, cfg, exp, seed=4, return_esteps=True); dt = time.perf_counter() - t0
    t0 = time.perf_counter(); mse, es = optimizer.objective_batched(P, cfg, exp, seed=5, return_esteps=True); dt = time.perf_counter() - t0
    res["c4_" + exp] = dict(workload=f"C4 {exp}: 4096 Sobol candidates in DEFAULT_BOUNDS, one mcl_objective call (wall clock incl. host tables + D2H)",
                            seconds=dt, esteps=es, esteps_per_s=es / dt, finite=int(np.isfinite(mse).sum()),
                            objectives_per_s=4096 / dt)
print(json.dumps(res, indent=1))
