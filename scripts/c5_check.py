"""C5 (TL ramp, 2000-electron boxes): pipelined vs in-order fingerprints for one CTA width.  usage: python scripts/c5_check.py <nt> [replicas]"""
import os, sys
import numpy as np
sys.path.insert(0, ".")
from mcluminescence_b200 import engine, workloads
nt, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 300
wl = workloads.c5(n_replicas=n)
os.environ["MCL_PHILOX_NT"] = nt
outs = []
for pipe in ("1", "0"):
    os.environ["MCL_PHILOX_PIPE"] = pipe
    outs.append(engine.run_replicas(wl["replicas"], wl["segments"], wl["max_steps"], seed=43, hist=wl["hist"], trace=False, sync=True))
a, b = outs
bad = [k for k in ("status", "steps_used", "final_n_e", "esteps", "hist_events", "hist_occ") if not np.array_equal(np.asarray(getattr(a, k)), np.asarray(getattr(b, k)))]
print(f"c5 nt={nt}: {'IDENTICAL' if not bad else 'DIFFERENT ' + str(bad)}  status {np.unique(np.asarray(a.status)).tolist()}")
