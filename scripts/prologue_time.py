"""How much of a C2 ensemble is the prologue (seeding + candidate lists)?  The same boxes with legs that end after their first step.
usage: python scripts/prologue_time.py [replicas]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
from mcluminescence_b200 import engine, workloads
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
for tag, dur in (("full legs", None), ("one step per leg", [1e-12, 1e-12])):
    wl = workloads.c2(n_replicas=n)
    if dur:
        wl["segments"]["duration"] = dur
    for i in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        o = engine.run_replicas(wl["replicas"], wl["segments"], wl["max_steps"], seed=7 + i, hist=wl["hist"], trace=False, sync=True)
        dt = time.perf_counter() - t0
    print(f"{tag}: {dt * 1e3:.1f} ms for {n} replicas, steps per replica {np.asarray(o.steps_used).mean():.1f}", flush=True)
