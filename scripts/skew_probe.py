"""Per-warp cycle accounting of the step loop (sweep / wait at the step barrier / rest), from a library built with
-DMCL_PROFILE_SKEW (scripts/build_variant.sh skew -DMCL_PROFILE_SKEW).  usage: python scripts/skew_probe.py c2 444"""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, ".")
os.environ.setdefault("MCL_B200_LIB", os.path.abspath("scripts/ab_libs/skew.so"))
from mcluminescence_b200 import ensemble, workloads
wl = workloads.c2(n_replicas=int(sys.argv[2])) if sys.argv[1] == "c2" else workloads.c5(n_replicas=int(sys.argv[2]))
lib = ctypes.CDLL(os.environ["MCL_B200_LIB"])
buf = (ctypes.c_ulonglong * 320)()
ensemble.run_ensemble(wl, seed=7)
lib.mcl_debug_prof(buf, 1)
ensemble.run_ensemble(wl, seed=8)
assert lib.mcl_debug_prof(buf, 1) == 0
a = np.array(buf, dtype=np.float64).reshape(10, 32)
nw = int((a[3] > 0).sum())
print(sys.argv[1], "warps", nw, "steps/warp", a[3, 0])
for w in range(nw):
    tot = a[0, w] + a[1, w] + a[2, w]
    print(f"warp {w}: per step cycles sweep {a[0, w] / a[3, w]:8.0f}  wait {a[1, w] / a[3, w]:8.0f}  rest {a[2, w] / a[3, w]:8.0f}   "
          f"shares {a[0, w] / tot:.3f} {a[1, w] / tot:.3f} {a[2, w] / tot:.3f}")
names = ["reduce+decision", "histogram", "retire pair", "scan+lists", "warp searches"]
print("grid searches that scanned ring R = 1, 2, ...:", [int(v) for v in a[9, 1:9]], "in", int(a[3, 0]), "replica-steps")
print("after the barrier, cycles per step (mean over warps): " + ", ".join(f"{n} {a[4 + i, :nw].sum() / a[3, :nw].sum():.0f}" for i, n in enumerate(names)))
