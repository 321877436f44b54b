"""GPU parity, replay mode: the CUDA FP64 kernel consumes the reference's uniform draws and must
reproduce the reference's integer traces bit-for-bit (golden vectors from the unmodified
reference) and agree with the CPU oracle on fresh seeded inputs.  Everything goes through the
C ABI (mcl_run / mcl_run_host)."""
import ctypes as C
import io
from contextlib import redirect_stdout

import numpy as np
import pytest

from tests import helpers

pytestmark = pytest.mark.gpu

T_RTOL = 1e-12          # CUDA exp/log vs NumPy's AVX512 / glibc: <= 1 ulp apart, never a different event
SIM_CASES = ["sim_kat1", "sim_kat2", "sim_sweep2x2", "sim_partial", "sim_empty", "sim_labphys",
             "sim_zero_duration", "sim_kat0"]


@pytest.fixture(scope="module")
def gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("marked gpu but no CUDA device is visible")
    from mcluminescence_b200 import _native
    _native.load()
    return torch


@pytest.mark.parametrize("name", SIM_CASES)
def test_simulate_replay_matches_reference(gpu, golden, name, tmp_path):
    from mcluminescence_b200 import simulate as sim_mod
    from mcluminescence_b200.config import compose
    meta, arrs = golden.meta(name), golden.arrays(name)
    cfg = compose(overrides=meta["overrides"])
    sim_mod.PROJECT_ROOT = str(tmp_path)
    x_ax, lum, er, configs = sim_mod.simulate(cfg, rng="replay", seed=meta["seed"])
    steps, sims, runs = lum.shape
    traces = helpers.split_traces(arrs)
    for run in range(runs):
        N_e = configs[run]["exp_type_fp"]["N_e"]
        for j in range(sims):
            ev, ne, tt = traces[run * sims + j]
            n = len(ev)
            assert np.count_nonzero(x_ax[:, j, run] > 0) == n
            assert np.array_equal(lum[:n, j, run], ev.astype(np.float64))
            assert np.array_equal(er[:n, j, run], ne / N_e)          # same ints -> same ratios
            np.testing.assert_allclose(x_ax[:n, j, run], tt, rtol=T_RTOL, atol=0)
            assert not lum[n:, j, run].any() and not x_ax[n:, j, run].any()
    import pandas as pd
    df = pd.read_csv(tmp_path / "results" / "simulations" / "exp_.csv")
    assert list(df.columns) == ["run", "sim", "step", "lum", "electron_ratio"]
    assert len(df) == steps * sims * runs
    assert int(df.lum.sum()) == int(sum(meta["events"]))


@pytest.mark.parametrize("name", [
    "lab_default_tl_clbr", "lab_default_tl_fsm13", "lab_default_iso",
    "lab_kat3_tl_clbr", "lab_kat3_tl_fsm13", "lab_kat3_iso",
    "lab_sobol0_tl_clbr", "lab_sobol1_tl_clbr", "lab_sobol2_tl_clbr", "lab_sobol3_tl_clbr",
    "lab_sobol0_iso", "lab_sobol1_iso"])
def test_objective_replay_matches_reference(gpu, golden, name):
    from mcluminescence_b200 import engine, optimizer
    from mcluminescence_b200.config import compose
    meta = golden.meta(name)
    cfg = compose(overrides=helpers.LAB_OVERRIDES)
    cfg["rng"] = "replay"
    engine.seed_replay(meta["seed"])
    buf = io.StringIO()
    with redirect_stdout(buf):
        if meta["p"] is None:
            val = optimizer.run_one_sim(cfg, meta["exp"])
        else:
            val = optimizer.objective(np.asarray(meta["p"]), cfg, meta["exp"])
    assert val == meta["value"]
    assert engine.global_replay().pos == meta["n_uniforms"]
    assert buf.getvalue().strip() == meta["printed"]


def test_structure_log_matches_oracle_on_fresh_inputs(gpu):
    """Fresh seeded configs (not in the golden set): GPU replay vs CPU oracle, event by event."""
    from mcluminescence_b200 import engine
    from mcluminescence_b200.config import compose, initialize_runs
    from mcluminescence_b200.replicas import simulate_tables
    from oracle import mcl_oracle as mo
    rs = np.random.RandomState(2024)
    for trial in range(6):
        N_e = int(rs.randint(20, 260))
        holes = int(rs.randint(N_e, 2 * N_e + 5))
        ov = [f"exp_type_fp.N_e={N_e}", f"exp_type_fp.holes={holes}", "exp_type_fp.sims=2",
              "exp_type_fp.steps=4000", f"exp_type_fp.T_rate=[{rs.choice([0.5, 2, 10, 20])}]",
              f"exp_type_fp.duration=[{rs.choice([30, 60, 200])}]",
              f"exp_type_fp.e_ratio_start={rs.choice([1, 0.5, 0.9])}",
              f"exp_type_fp.boundary_factor=[{rs.choice([1.0, 1.2, 1.5])}]",
              f"physics_fp.E_loc_2={rs.choice([1.2, 1.1])}", f"physics_fp.Retrap={rs.choice([0.5, 0.2])}"]
        if trial % 2:
            ov = ["physics_fp=lab_TL", "exp_type_fp.rho_prime=1e-5"] + ov[:-2]
        cfg = compose(overrides=ov)
        runs = initialize_runs(cfg)
        reps, segs = simulate_tables(runs, 2)
        seed = 1000 + trial
        ref = mo.run(reps, segs, 4000, seed=seed)
        assert ref.rc == 0
        got = engine.run_replay_chained(reps, segs, 4000, engine.ReplayStream(seed), structure=True)
        assert np.array_equal(got["steps_used"], ref.steps_used), ov
        assert np.array_equal(got["consumed"], ref.consumed)
        assert np.array_equal(got["esteps"], ref.esteps)
        for r in range(len(reps)):
            n = int(ref.steps_used[r])
            for k in ("event", "n_e", "kind", "e_idx", "h_idx"):
                assert np.array_equal(got[k][r, :n], getattr(ref, k)[r, :n]), (ov, r, k)
            np.testing.assert_allclose(got["t"][r, :n], ref.t[r, :n], rtol=T_RTOL, atol=0)


def test_run_host_entry_point_and_errors(gpu, golden):
    """mcl_run_host: plain host buffers in, host buffers out (what a non-torch caller binds)."""
    from mcluminescence_b200 import _native
    from oracle import mcl_oracle as mo
    L = _native.load()
    meta = golden.meta("sim_kat2")
    reps, segs, steps, _, _ = helpers.sim_tables_for(meta)
    ref = mo.run(reps, segs, steps, seed=meta["seed"])
    n_u = int(ref.consumed.sum())
    u = np.random.RandomState(meta["seed"]).random_sample(n_u)
    off = np.array([0, n_u], dtype=np.int64)
    ev = np.zeros((1, steps), np.int32); ne = np.zeros((1, steps), np.int32); tt = np.zeros((1, steps))
    used = np.zeros(1, np.int32); status = np.zeros(1, np.int32); consumed = np.zeros(1, np.int64)
    a = _native.RunArgs()
    a.replicas, a.n_replicas = reps.ctypes.data, 1
    a.segments, a.n_segments = segs.ctypes.data, len(segs)
    a.max_steps, a.mode = steps, 1
    a.replay_u, a.replay_off = u.ctypes.data, off.ctypes.data
    a.event, a.n_e, a.t = ev.ctypes.data, ne.ctypes.data, tt.ctypes.data
    a.steps_used, a.status, a.consumed = used.ctypes.data, status.ctypes.data, consumed.ctypes.data
    assert L.mcl_run_host(C.byref(a)) == 0, _native.last_error()
    n = int(ref.steps_used[0])
    assert used[0] == n and status[0] == 0 and consumed[0] == n_u
    assert np.array_equal(ev[0, :n], ref.event[0, :n]) and np.array_equal(ne[0, :n], ref.n_e[0, :n])
    # stream too short -> per-replica status MCL_ERR_STREAM, call itself succeeds
    off2 = np.array([0, n_u - 10], dtype=np.int64)
    a.replay_off = off2.ctypes.data
    assert L.mcl_run_host(C.byref(a)) == 0
    assert status[0] == -3
    # record capacity too small -> MCL_ERR_STEPS (the reference raises IndexError)
    a.replay_off = off.ctypes.data
    a.max_steps = 10
    assert L.mcl_run_host(C.byref(a)) == 0
    assert status[0] == -1
    # bad arguments are rejected before any launch
    a.max_steps = 0
    assert L.mcl_run_host(C.byref(a)) == -6 and "max_steps" in _native.last_error()


def test_simulate_raises_index_error_like_reference(gpu, golden, tmp_path):
    from mcluminescence_b200 import simulate as sim_mod
    from mcluminescence_b200.config import compose
    meta = golden.meta("sim_kat1")
    cfg = compose(overrides=meta["overrides"] + ["exp_type_fp.steps=50"])
    sim_mod.PROJECT_ROOT = str(tmp_path)
    with pytest.raises(IndexError):
        sim_mod.simulate(cfg, rng="replay", seed=meta["seed"])
    with pytest.raises(IndexError):
        sim_mod.simulate(cfg, rng="philox", seed=3)
