"""Record-for-record comparison of two builds of the library on a set of workloads (each build runs in its own process).

    python scripts/compare_libs.py <libA.so> <libB.so>      # exit code 0 = identical
"""
import os, subprocess, sys, tempfile
import numpy as np

def dump(out):
    sys.path.insert(0, ".")
    from mcluminescence_b200 import engine, workloads
    from mcluminescence_b200.config import compose, initialize_runs
    from mcluminescence_b200.replicas import LAB_CSV, LabTable
    from mcluminescence_b200.tl_trap_lab import PROJECT_ROOT
    res = {}
    def put(tag, o):
        o.raise_on_error()
        for k in ("event", "n_e", "t", "steps_used", "final_n_e", "esteps", "hist_events", "hist_occ", "obs_n_e"):
            v = getattr(o, k, None)
            if v is not None:
                res[f"{tag}.{k}"] = np.asarray(v)
    for tag, wl in (("c2", workloads.c2(n_replicas=8)), ("c5", workloads.c5(n_replicas=200)), ("c2s", workloads.c2(n_replicas=40, n_e=600, n_bins=40)),
                    ("c2two", workloads.c2(n_replicas=6, n_e=5000, physics_overrides=["physics_fp.E_loc_2=1.0", "physics_fp.Retrap=0.3"]))):
        put(tag, engine.run_replicas(wl["replicas"], wl["segments"], wl["max_steps"], seed=5, hist=wl["hist"], trace=True, sync=True))
    os.environ["MCL_PHILOX_NT"] = "256"            # the production width of 10^4-electron boxes (few replicas would get 512)
    wl = workloads.c2(n_replicas=6)
    put("c2nt256", engine.run_replicas(wl["replicas"], wl["segments"], wl["max_steps"], seed=5, hist=wl["hist"], trace=True, sync=True))
    os.environ["MCL_PHILOX_SHARE_BM"] = "2"       # + self-check of the scan-skip masks (any miss = MCL_ERR_INTERNAL)
    put("c2nt256v", engine.run_replicas(wl["replicas"], wl["segments"], wl["max_steps"], seed=5, hist=wl["hist"], trace=True, sync=True))
    del os.environ["MCL_PHILOX_NT"], os.environ["MCL_PHILOX_SHARE_BM"]
    run = initialize_runs(compose(overrides=["exp_type_fp=TLlab", "physics_fp=lab_TL"]))[0]
    for exp in ("tl_clbr", "iso"):
        lt = LabTable(*LAB_CSV[exp], PROJECT_ROOT)
        r, s = lt.tables(run)
        put(exp, engine.run_replicas(r, s, 20000, seed=6, obs_time=lt.obs_time, trace=True, sync=True))
    np.savez(out, **res)

if sys.argv[1] == "--dump":
    dump(sys.argv[2])
else:
    outs = []
    for lib in sys.argv[1:3]:
        f = tempfile.mktemp(suffix=".npz")
        subprocess.run([sys.executable, __file__, "--dump", f], check=True, env=dict(os.environ, MCL_B200_LIB=os.path.abspath(lib)))
        outs.append(np.load(f))
    bad = [k for k in outs[0].files if not np.array_equal(outs[0][k], outs[1][k])]
    print("compared", len(outs[0].files), "arrays;", "IDENTICAL" if not bad else f"DIFFERENT: {bad}")
    sys.exit(1 if bad else 0)
