import sys, numpy as np
sys.path.insert(0, ".")
from mcluminescence_b200 import optimizer, workloads
from mcluminescence_b200.config import compose
cfg = compose(overrides=["exp_type_fp=TLlab", "physics_fp=lab_TL"])
P = workloads.c4_candidates(2048, seed=4)
for s in (1, 2):
    mse, es = optimizer.objective_batched(P, cfg, "tl_clbr", seed=s, return_esteps=True)
print(es)
