"""GPU: the callers either side of the hot path: the Optimizer train / replay driver with the
reference's CSV format, and the glow-curve post-processing against the reference's `hist_and_smooth` definition."""
import numpy as np
import pytest

from tests import helpers

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("marked gpu but no CUDA device is visible")
    from mcluminescence_b200 import _native
    _native.load()
    return torch


def test_optimizer_main_train_appends_and_replay_reads_best_row(gpu, tmp_path, capsys):
    """reference optimizer.py:104-116 (differential evolution, one CSV row `param_0..9,mse` appended per run, header
    only for a new file) and :124-133 (`task=replay` re-simulates the row with the smallest mse)."""
    import pandas as pd
    from mcluminescence_b200 import optimizer
    from mcluminescence_b200.config import compose
    optimizer.PROJECT_ROOT = str(tmp_path)
    try:
        base = helpers.LAB_OVERRIDES + ["+exp=tl_clbr", "+gens=1", "+pop=2", "+seed=3"]
        optimizer.main(compose("config_fp", base + ["+task=train"]))
        csv = tmp_path / "results" / "lab_sims" / "result_tl_clbr.csv"
        assert csv.exists()
        df = pd.read_csv(csv)
        assert list(df.columns) == [f"param_{i}" for i in range(10)] + ["mse"] and len(df) == 1
        lo, hi = np.array(optimizer.DEFAULT_BOUNDS).T
        p = df.iloc[0].filter(like="param_").values.astype(float)
        assert np.all(p >= lo) and np.all(p <= hi) and 0.0 <= df.mse[0] < 1.0
        out = capsys.readouterr().out
        assert "Saved run to" in out and "absError=" in out              # the best candidate is re-simulated and printed
        # a second run appends WITHOUT a header
        optimizer.main(compose("config_fp", base + ["+task=train", "+seed=4"]))
        text = csv.read_text().strip().splitlines()
        assert len(text) == 3 and text[0].startswith("param_0,") and not text[2].startswith("param_")
        df = pd.read_csv(csv)
        assert len(df) == 2
        capsys.readouterr()
        # replay: the row with the smallest mse is the one that is re-simulated
        optimizer.main(compose("config_fp", base + ["+task=replay"]))
        out = capsys.readouterr().out
        best = df.loc[df.mse.idxmin()]
        assert "Re-simulated best parameters for exp: tl_clbr" in out
        assert f"rho'={float(best.param_0)}" in out and f"P_retrap={float(best.param_9)}" in out
        with pytest.raises(ValueError):
            optimizer.main(compose("config_fp", base + ["+task=nope"]))
        with pytest.raises(FileNotFoundError):
            optimizer.main(compose("config_fp", helpers.LAB_OVERRIDES + ["+exp=iso", "+task=replay"]))
    finally:
        optimizer.PROJECT_ROOT = None


def reference_hist_and_smooth(t_axis, events, bin_width=1.0, win_deg=50.0):
    """The DEFINITION of the reference's glow-curve post-processing (src/class/plots.py:39-47 with running_mean at
    :19-21), restated for the test: 1 degC bins from 0 to max(T) + 1, events as weights, boxcar mean over 50 degC."""
    bins = np.arange(0, t_axis.max() + bin_width, bin_width)
    hist = np.histogram(t_axis, bins=bins, weights=events)[0] / bin_width
    k = max(1, int(win_deg / bin_width))
    return np.convolve(hist, np.ones(k) / k, "valid")


def test_kernel_temperature_histogram_gives_the_reference_glow_curve(gpu):
    """The kernel's fused MCL_AXIS_TEMP histogram + `postprocess.glow_curve` == `hist_and_smooth` applied to the same
    replica's (temperature, event) trace, which is what reference plots.py:57-66 feeds it for one replica."""
    from mcluminescence_b200 import engine, postprocess, workloads
    from mcluminescence_b200.engine import AXIS_TEMP, HistSpec
    wl = workloads.c1()                                   # the shipped default config: 4 heating rates x 2 sims
    reps, segs = wl["replicas"], wl["segments"]
    n_bins = 800
    hist = HistSpec(axis=AXIS_TEMP, n_bins=n_bins, lo=0.0, hi=float(n_bins), n_groups=len(reps))
    out = engine.run_replicas(reps, segs, wl["max_steps"], seed=77, hist=hist, hist_group=np.arange(len(reps), dtype=np.int32),
                              trace=True, sync=True)
    out.raise_on_error()
    for r in range(len(reps)):
        seg = segs[int(reps["seg_begin"][r])]
        n = int(out.steps_used[r])
        T_axis = float(seg["T_start"]) + float(seg["T_rate"]) * out.t[r, :n]           # plots.py:63-64
        events = out.event[r, :n].astype(float)
        want = reference_hist_and_smooth(T_axis, events)
        got = postprocess.glow_curve(out.hist_events[r], n_replicas=1)
        m = min(len(want), len(got))                      # the reference's axis ends at max(T) (just above 800 degC)
        assert m >= 700
        assert np.allclose(got[:m], want[:m], rtol=0, atol=1e-12), (r, np.abs(got[:m] - want[:m]).max())
        assert want[:m].max() > 0


def test_legacy_est_params_semantics_on_the_gpu(gpu, capsys):
    """the pre-refactor TL code (reference src/est_params/functions.py:270-360), selectable with
    `legacy=True`.  GPU (native Philox) against the oracle's legacy mode, which is pinned bit-for-bit against the
    unmodified legacy function (tests/test_oracle_golden.py); batched == single; and the reference's recorded optimum
    (results/lab_sims/result_tl_clbr.csv, first row) scores a few 1e-3 under its own semantics but ~0.02 under the
    current ones."""
    from mcluminescence_b200 import engine, optimizer
    from mcluminescence_b200.config import compose
    from oracle import mcl_oracle as mo
    golden = helpers.Golden()
    cfg = compose(overrides=helpers.LAB_OVERRIDES)
    for name in ("legacy_default", "legacy_best_row", "legacy_sobol2"):
        meta = golden.meta(name)
        lt, reps1, segs, run = helpers.legacy_setup(meta)
        assert np.all(reps1["protocol"] == 3)
        M, n_rows, steps = 96, len(reps1), int(run["exp_type_fp"]["steps"])
        reps = np.tile(reps1, M)
        out = engine.run_replicas(reps, segs, steps, seed=700, trace=False, sync=True)
        out.raise_on_error()
        ref = mo.run(reps, segs, steps, seed=800, parallel=True, trace=False)
        assert ref.rc == 0
        g, o = out.final_n_e.reshape(M, n_rows).astype(float), ref.final_n_e.reshape(M, n_rows).astype(float)
        for k in range(n_rows):
            se = np.sqrt(g[:, k].var(ddof=1) / M + o[:, k].var(ddof=1) / M)
            assert abs(g[:, k].mean() - o[:, k].mean()) <= 4.0 * se + 0.3, (name, k, g[:, k].mean(), o[:, k].mean(), se)
        ge, oe = out.esteps.reshape(M, n_rows).sum(1).astype(float), ref.esteps.reshape(M, n_rows).sum(1).astype(float)
        se = np.sqrt(ge.var(ddof=1) / M + oe.var(ddof=1) / M)
        assert abs(ge.mean() - oe.mean()) <= 4.0 * se, (name, ge.mean(), oe.mean(), se)
    # the Optimizer seam: batched == single candidate, legacy != current
    best = np.asarray(golden.meta("legacy_best_row")["p"])
    P = np.tile(best[:, None], (1, 64))
    leg = optimizer.objective_batched(P, cfg, "tl_clbr", seed=9, legacy=True)
    cur = optimizer.objective_batched(P, cfg, "tl_clbr", seed=9)
    one = optimizer.objective(best, cfg, "tl_clbr", seed=9, candidate_id=5, legacy=True)
    assert one == leg[5]
    assert leg.mean() < 0.008 and leg.min() < 0.004 and cur.mean() > 2.0 * leg.mean(), (leg.mean(), leg.min(), cur.mean())
    with pytest.raises(Exception):
        optimizer.objective_batched(P, cfg, "iso", seed=9, legacy=True)
    capsys.readouterr()
