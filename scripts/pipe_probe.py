"""Counters and per-phase cycles of the pipelined step loop, from a library built with -DMCL_PIPE_STATS
(scripts/build_variant.sh pstat -DMCL_PIPE_STATS).  usage: [MCL_PHILOX_NT=..] python scripts/pipe_probe.py 444 [c2|c5]"""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, ".")
os.environ.setdefault("MCL_B200_LIB", os.path.abspath("scripts/ab_libs/pstat.so"))
from mcluminescence_b200 import ensemble, workloads
wl = workloads.c5(n_replicas=int(sys.argv[1])) if sys.argv[2:] == ["c5"] else workloads.c2(n_replicas=int(sys.argv[1]))
lib = ctypes.CDLL(os.environ["MCL_B200_LIB"])
buf = (ctypes.c_ulonglong * 64)()
ensemble.run_ensemble(wl, seed=7)
lib.mcl_debug_pipe_stats(buf, 1)
ensemble.run_ensemble(wl, seed=8)
assert lib.mcl_debug_pipe_stats(buf, 1) == 0
full = np.array(buf, dtype=np.float64)
a, hh = full[:32], full[32:].reshape(4, 8)
steps = a[0]
print(f"steps {steps:.0f}  entries {a[1]:.0f} ({steps / max(a[1], 1):.0f} steps per entry)  hand-backs: leg end {a[2]:.0f}, compaction {a[3]:.0f}, fill clock {a[4]:.0f}, overflow {a[5]:.0f}, conduction band {a[7]:.0f}")
print(f"per step: thread re-evaluations {a[6] / steps:.4f}, stale winners re-targeted {a[8] / steps:.4f}, grid searches {a[9] / steps:.4f}")
print(f"sweep team (warp 0), cycles per step: sweep {a[16] / steps:.0f}, wait DONE {a[17] / steps:.0f}, entry + arrive {a[18] / steps:.0f}")
print(f"decision warp, cycles per step: wait FULL {a[20] / steps:.0f}, minimum + re-evaluations {a[22] / steps:.0f}, "
      f"decision + event {a[23] / steps:.0f}, histogram {a[24] / steps:.0f}")
print("sweep cycles per step, team warps 0..6: " + ", ".join(f"{a[25 + w] / steps:.0f}" for w in range(7)))
edges = ["<3000", "<4500", "<6000", "<8000", "<12000", "<20000", "<40000", ">=40000"]
print("decision-warp work per step (FULL -> end of histogram), share of steps / of its cycles per bucket:")
print("   " + "  ".join(f"{e}: {hh[0, i] / max(hh[0].sum(), 1):.3f}/{hh[1, i] / max(hh[1].sum(), 1):.3f}" for i, e in enumerate(edges)))
print("warp 0 waiting for the flag, share of waits / of waited cycles per bucket (cycles per step: %.0f):" % (hh[3].sum() / steps))
print("   " + "  ".join(f"{e}: {hh[2, i] / max(hh[2].sum(), 1):.3f}/{hh[3, i] / max(hh[3].sum(), 1):.3f}" for i, e in enumerate(edges)))
