"""Small invocations of every kernel path for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from mcluminescence_b200 import engine, workloads
from mcluminescence_b200.config import compose, initialize_runs
from mcluminescence_b200.replicas import LAB_CSV, LabTable, simulate_tables
from mcluminescence_b200.tl_trap_lab import PROJECT_ROOT

cfg = compose(overrides=["exp_type_fp.N_e=120", "exp_type_fp.holes=150", "exp_type_fp.steps=1500",
                         "exp_type_fp.T_rate=[20]", "exp_type_fp.duration=[30]", "exp_type_fp.sims=3"])
reps, segs = simulate_tables(initialize_runs(cfg), 3)
out = engine.run_replicas(reps, segs, 1500, seed=1, sync=True); out.raise_on_error()       # NT=32 path
got = engine.run_replay_chained(reps[:1], segs, 1500, engine.ReplayStream(3))               # replay kernel
wl = workloads.c2(n_replicas=4, n_e=600, n_bins=32)                                         # NT=64, two legs, histograms
wl["segments"]["duration"] = [100.0, 200.0]
out = engine.run_replicas(wl["replicas"], wl["segments"], wl["max_steps"], seed=2, hist=wl["hist"], sync=True); out.raise_on_error()
wl = workloads.c2(n_replicas=2, n_e=5000, n_bins=32)                                        # NT=256, PPC=2
wl["segments"]["duration"] = [50.0, 20.0]
out = engine.run_replicas(wl["replicas"], wl["segments"], wl["max_steps"], seed=2, hist=wl["hist"], trace=False, sync=True); out.raise_on_error()
run = initialize_runs(compose(overrides=["exp_type_fp=TLlab", "physics_fp=lab_TL"]))[0]      # fills, lab protocols
for exp in ("tl_clbr", "iso"):
    lt = LabTable(*LAB_CSV[exp], PROJECT_ROOT)
    r, s = lt.tables(run)
    out = engine.run_replicas(r, s, 20000, seed=4, obs_time=lt.obs_time, trace=False, sync=True); out.raise_on_error()
    if exp == "tl_clbr":                                                                     # legacy semantics (one-warp kernel)
        r, s = lt.tables(run, legacy=True)
        out = engine.run_replicas(r, s, 20000, seed=4, trace=True, sync=True); out.raise_on_error()
import os
os.environ["MCL_SMALLBOX"] = "0"                                                            # the same rows on the block kernel: fills,
os.environ["MCL_PHILOX_FILL_EXTRA"] = "-92"                                                 # regrids, shared-memory slab
for exp in ("tl_clbr", "iso"):
    lt = LabTable(*LAB_CSV[exp], PROJECT_ROOT)
    r, s = lt.tables(run)
    out = engine.run_replicas(r, s, 20000, seed=4, obs_time=lt.obs_time, trace=False, sync=True); out.raise_on_error()
del os.environ["MCL_SMALLBOX"], os.environ["MCL_PHILOX_FILL_EXTRA"]
wl = workloads.c3(replicas_per_dose=1)                                                      # dose -> TL on one box: regrid + list rebuild
wl["replicas"]["N_e"] = 300
out = engine.run_replicas(wl["replicas"][:3], wl["segments"], wl["max_steps"], seed=5, hist=wl["hist"], hist_group=wl["hist_group"][:3],
                          trace=False, sync=True); out.raise_on_error()
print("sanitize_small ok", int(out.esteps.sum()))
