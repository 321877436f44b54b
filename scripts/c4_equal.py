"""Two builds of the library on Optimizer populations (one-warp kernel): objective values and electron-step counts must be equal.
usage: python scripts/c4_equal.py <libA.so> <libB.so>"""
import os, subprocess, sys, tempfile
import numpy as np

def dump(out):
    sys.path.insert(0, ".")
    from mcluminescence_b200 import optimizer, workloads
    from mcluminescence_b200.config import compose
    cfg = compose(overrides=["exp_type_fp=TLlab", "physics_fp=lab_TL"])
    P = workloads.c4_candidates(1024, seed=11)
    res = {}
    for exp in ("tl_clbr", "iso"):
        for legacy in ((False, True) if exp == "tl_clbr" else (False,)):
            mse, es = optimizer.objective_batched(P, cfg, exp, seed=77, return_esteps=True, legacy=legacy)
            res[f"{exp}.{legacy}.mse"] = mse; res[f"{exp}.{legacy}.es"] = np.array([es])
    np.savez(out, **res)

if sys.argv[1] == "--dump":
    dump(sys.argv[2])
else:
    outs = []
    for lib in sys.argv[1:3]:
        f = tempfile.mktemp(suffix=".npz")
        subprocess.run([sys.executable, __file__, "--dump", f], check=True, env=dict(os.environ, MCL_B200_LIB=os.path.abspath(lib)))
        outs.append(np.load(f))
    bad = [k for k in outs[0].files if not np.array_equal(outs[0][k], outs[1][k], equal_nan=True)]
    print("compared", len(outs[0].files), "arrays;", "IDENTICAL" if not bad else f"DIFFERENT: {bad}")
    sys.exit(1 if bad else 0)
