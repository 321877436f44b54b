"""Wall-clock breakdown of Optimizer populations (mcl_objective phases + the Python around it)."""
import os, sys, time
sys.path.insert(0, ".")
os.environ["MCL_OBJECTIVE_TIMING"] = "1"
import numpy as np, torch
from mcluminescence_b200 import optimizer, workloads
from mcluminescence_b200.config import compose
cfg = compose(overrides=["exp_type_fp=TLlab", "physics_fp=lab_TL"])
for S in (4096, 4096, 4096, 16384, 16384, 256, 256):
    P = workloads.c4_candidates(S, seed=4)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    mse, es = optimizer.objective_batched(P, cfg, "tl_clbr", seed=5, return_esteps=True)
    dt = time.perf_counter() - t0
    print(f"S={S}: python wall {dt*1e3:.2f} ms, {es/dt/1e9:.2f}e9 e-steps/s, finite {int(np.isfinite(mse).sum())}", file=sys.stderr)
