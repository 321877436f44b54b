"""usage: python scripts/c5_time.py <nt> <replicas>  -- one C5 ensemble, wall clock"""
import os, sys, time
import numpy as np
sys.path.insert(0, ".")
from mcluminescence_b200 import engine, workloads
os.environ["MCL_PHILOX_NT"] = sys.argv[1]
wl = workloads.c5(n_replicas=int(sys.argv[2]))
for i in range(2):
    t0 = time.perf_counter()
    o = engine.run_replicas(wl["replicas"], wl["segments"], wl["max_steps"], seed=43 + i, hist=wl["hist"], trace=False, sync=True)
    print(f"c5 nt={sys.argv[1]} R={sys.argv[2]}: {time.perf_counter() - t0:.3f} s, status {np.unique(np.asarray(o.status)).tolist()}, steps max {int(np.asarray(o.steps_used).max())}", flush=True)
