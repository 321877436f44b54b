"""Time the UNMODIFIED NumPy reference (``oracle/_ref/src_class``, see make_ref.py) on this host: one core, a
down-scaled isothermal hold leg of the C2 schedule through the reference's own ``simulate(cfg)``.

MEASUREMENT INFRASTRUCTURE ONLY (bench.py's `cpu_baseline.numpy_reference`).  The reference cannot run the C2 shape
itself: it keeps a dense n_e x n_h float64 distance matrix (1.4 GB at 10^4 x 17 279, 6 `np.delete` copies per event) and
has neither an optical leg nor multi-leg schedules.  So the sample is the hold leg (250 degC, no ramp) at N_e = holes =
`n_e`, sized to the time budget.  Electron-steps are counted from the reference's own outputs, as the survey defines
them: sum over steps of n_e before the event.

    python oracle/ref_harness/time_ref.py [seconds_budget]   ->  one JSON line
"""
from __future__ import annotations

import io
import json
import os
import sys
import tempfile
import time
from contextlib import redirect_stdout

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.abspath(os.path.join(HERE, "..", ".."))
REF = os.path.join(REPO, "oracle", "_ref", "src_class")


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 15.0
    for p in (REF, os.path.join(HERE, "shims"), REPO):
        if p not in sys.path:
            sys.path.insert(0, p)
    import numpy as np
    import simulate  # the reference's own module  # type: ignore
    from mcluminescence_b200.config import compose      # the config surface only (same keys and values as the reference's conf/)
    out_root = tempfile.mkdtemp(prefix="mcl_ref_time_")
    os.makedirs(os.path.join(out_root, "results", "simulations"), exist_ok=True)
    simulate.PROJECT_ROOT = out_root

    def run(n_e, duration):
        cfg = compose("config_fp", [f"exp_type_fp.N_e={n_e}", f"exp_type_fp.holes={n_e}", "exp_type_fp.T_start=[250]",
                                    "exp_type_fp.T_rate=[0]", f"exp_type_fp.duration=[{duration}]", "exp_type_fp.sims=1",
                                    f"exp_type_fp.steps={4 * n_e}"])
        np.random.seed(7)
        t0 = time.perf_counter()
        with redirect_stdout(io.StringIO()):
            x_ax, lum, er, _ = simulate.simulate(cfg)
        dt = time.perf_counter() - t0
        used = int(np.count_nonzero(x_ax[:, 0, 0] > 0))
        n_after = np.rint(er[:used, 0, 0] * n_e)
        esteps = int((n_after + lum[:used, 0, 0]).sum())             # no dose: n_before = n_after + event
        return esteps, used, dt

    # a short probe sets the size of the real sample
    es, steps, dt = run(400, 1000.0)
    n_e = 1000 if dt * 12 < budget else 400
    es, steps, dt = run(n_e, 1000.0)
    total_es, total_dt, runs = es, dt, 1
    while total_dt + dt < budget:
        es, steps, dt = run(n_e, 1000.0)
        total_es += es; total_dt += dt; runs += 1
    print(json.dumps({"value": total_es / total_dt, "unit": "electron-steps/s", "cores": 1, "kind": "reference",
                      "sample": f"unmodified NumPy reference simulate(cfg): C2 hold leg (250 degC, 1000 s) at N_e = holes = {n_e}, "
                                f"{runs} runs, {total_dt:.1f} s on one core (the reference is single-threaded)",
                      "electron_steps": total_es}))


if __name__ == "__main__":
    main()
