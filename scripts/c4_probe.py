import sys, time, numpy as np, torch
sys.path.insert(0, ".")
from mcluminescence_b200 import optimizer, workloads
from mcluminescence_b200.config import compose
cfg = compose(overrides=["exp_type_fp=TLlab", "physics_fp=lab_TL"])
for S in (256, 1024, 4096, 16384):
    P = workloads.c4_candidates(S, seed=4)
    optimizer.objective_batched(P, cfg, "tl_clbr", seed=1)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    mse, es = optimizer.objective_batched(P, cfg, "tl_clbr", seed=2, return_esteps=True)
    dt = time.perf_counter() - t0
    print(f"S={S:6d} wall {dt*1e3:8.2f} ms  esteps {es:.3e}  {es/dt:.3e} e-steps/s  {S/dt:.0f} objectives/s")
