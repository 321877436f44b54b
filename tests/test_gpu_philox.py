"""GPU parity, native Philox mode: ensemble statistics of the FP32/SFU kernel against the CPU
oracle (float64, MT19937) on the same configurations.

Tolerances (north_star): ensemble means within 3 sigma of the replica spread -- here the combined
standard error of the two ensembles, with 4 sigma as the hard limit because several quantities
are checked per case -- plus a two-sample Kolmogorov-Smirnov test on the pooled event times."""
import numpy as np
import pytest

from tests import helpers

pytestmark = pytest.mark.gpu

R_ENS = 384
SIGMA = 4.0
KS_P = 1e-3


@pytest.fixture(scope="module")
def gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("marked gpu but no CUDA device is visible")
    from mcluminescence_b200 import _native
    _native.load()
    return torch


def ensemble_tables(overrides, R):
    from mcluminescence_b200.config import compose, initialize_runs
    from mcluminescence_b200.replicas import simulate_tables
    cfg = compose(overrides=overrides)
    runs = initialize_runs(cfg)
    assert len(runs) == 1
    reps, segs = simulate_tables(runs, R)
    return reps, segs, int(runs[0]["exp_type_fp"]["steps"])


def stats(event, n_e, t, used, grid, n_e0):
    """per-replica: events, final n_e, n_e sampled on a time grid; pooled event times."""
    R = event.shape[0]
    n_ev = np.array([event[r, :used[r]].sum() for r in range(R)], dtype=np.float64)
    fin = np.array([n_e[r, used[r] - 1] if used[r] else 0 for r in range(R)], dtype=np.float64)
    occ = np.zeros((R, len(grid)))
    times = []
    for r in range(R):
        tt, ne = t[r, :used[r]], n_e[r, :used[r]]
        idx = np.searchsorted(tt, grid, side="right") - 1        # last step at or before the grid time
        occ[r] = np.where(idx >= 0, ne[np.maximum(idx, 0)], n_e0)
        times.append(tt[event[r, :used[r]] > 0])
    return n_ev, fin, occ, np.concatenate(times)


def assert_means_agree(a, b, what):
    se = np.sqrt(a.var(ddof=1) / len(a) + b.var(ddof=1) / len(b))
    diff = abs(a.mean() - b.mean())
    assert diff <= SIGMA * se + 1e-12, f"{what}: GPU {a.mean():.4f} vs oracle {b.mean():.4f}, {diff / max(se, 1e-30):.2f} sigma"


CASES = {
    # TL ramp, tunnelling only (the default physics), heavy recombination
    "tl_ramp": (["exp_type_fp.N_e=200", "exp_type_fp.holes=200", "exp_type_fp.steps=2000",
                 "exp_type_fp.T_rate=[20]", "exp_type_fp.duration=[40]"], np.linspace(5, 39, 12)),
    # isothermal, two tunnelling channels
    "iso_two_channel": (["exp_type_fp.N_e=200", "exp_type_fp.holes=200", "exp_type_fp.steps=2000",
                         "exp_type_fp.T_start=[250]", "exp_type_fp.T_rate=[0]", "exp_type_fp.duration=[1000]",
                         "physics_fp.E_loc_2=1.0", "physics_fp.Retrap=0.3"], np.geomspace(1, 900, 12)),
    # conduction-band channel active (finite E_cb), partially filled
    "cb_channel": (["physics_fp=lab_TL", "exp_type_fp.N_e=120", "exp_type_fp.holes=150", "exp_type_fp.steps=3000",
                    "exp_type_fp.T_rate=[2]", "exp_type_fp.duration=[250]", "exp_type_fp.rho_prime=1e-5",
                    "exp_type_fp.e_ratio_start=0.8"], np.linspace(20, 240, 12)),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_simulate_ensemble_matches_oracle(gpu, name):
    from mcluminescence_b200 import engine
    from oracle import mcl_oracle as mo
    overrides, grid = CASES[name]
    reps, segs, steps = ensemble_tables(overrides, R_ENS)
    out = engine.run_replicas(reps, segs, steps, seed=20240 + len(name), sync=True)
    out.raise_on_error()
    ref = mo.run(reps, segs, steps, seed=77, parallel=True)
    assert ref.rc == 0
    g = stats(out.event, out.n_e, out.t, out.steps_used, grid, int(reps['n_e0'][0]))
    o = stats(ref.event, ref.n_e, ref.t, ref.steps_used, grid, int(reps['n_e0'][0]))
    assert_means_agree(g[0], o[0], f"{name}: events per replica")
    assert_means_agree(g[1], o[1], f"{name}: final n_e")
    for k in range(len(grid)):
        assert_means_agree(g[2][:, k], o[2][:, k], f"{name}: n_e(t={grid[k]:.3g})")
    from scipy.stats import ks_2samp
    ks = ks_2samp(g[3], o[3])
    assert ks.pvalue > KS_P, f"{name}: KS on event times p={ks.pvalue:.2e} (D={ks.statistic:.4f})"
    # bookkeeping identities of the kernel's own outputs
    est = np.array([(out.n_e[r, :out.steps_used[r]] + out.event[r, :out.steps_used[r]]).sum()
                    for r in range(R_ENS)])
    assert np.array_equal(est, out.esteps)              # no fills here: n_before = n_after + event


def test_results_depend_only_on_seed_and_global_replica_id(gpu):
    """Sharding invariance: replica g gives the same trace whatever batch / offset it ran in."""
    from mcluminescence_b200 import engine
    reps, segs, steps = ensemble_tables(CASES["tl_ramp"][0], 24)
    full = engine.run_replicas(reps, segs, steps, seed=5, sync=True)
    part = engine.run_replicas(reps[8:16], segs, steps, seed=5, replica_id0=8, sync=True)
    assert np.array_equal(full.event[8:16], part.event)
    assert np.array_equal(full.n_e[8:16], part.n_e)
    assert np.array_equal(full.t[8:16], part.t)
    other = engine.run_replicas(reps[8:16], segs, steps, seed=6, replica_id0=8, sync=True)
    assert not np.array_equal(other.event, part.event)


BLOCK = {"MCL_SMALLBOX": "0"}           # lab rows on the block kernel instead of the one-warp-per-replica kernel


@pytest.mark.parametrize("exp,knobs", [("tl_clbr", {}), ("iso", {}), ("tl_fsm-13", {}),
                                       ("tl_clbr", BLOCK), ("iso", BLOCK),
                                       # block kernel, regrid as soon as 8 fill-region slots are in use (default: N_e + 64): dozens of regrids per row
                                       ("tl_clbr", dict(BLOCK, MCL_PHILOX_FILL_EXTRA="-92")), ("iso", dict(BLOCK, MCL_PHILOX_FILL_EXTRA="-92")),
                                       # block kernel, hole tables in the HBM slab instead of shared memory
                                       ("tl_clbr", dict(BLOCK, MCL_PHILOX_SMEM_SLAB="0"))])
def test_lab_protocol_ensemble_matches_oracle(gpu, exp, knobs, monkeypatch):
    """TL_lab / ISO_lab with fills, stale caches and the conduction-band channel."""
    from mcluminescence_b200 import engine
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    from mcluminescence_b200.config import compose, initialize_runs
    from mcluminescence_b200.replicas import LAB_CSV, LabTable
    from oracle import mcl_oracle as mo
    cfg = compose(overrides=helpers.LAB_OVERRIDES)
    run = initialize_runs(cfg)[0]
    csv, proto = LAB_CSV[exp]
    lt = LabTable(csv, proto, helpers.DATA_ROOT)
    reps1, segs = lt.tables(run)
    n_rows, M = len(reps1), 96
    reps = np.tile(reps1, M)
    obs_time = np.tile(lt.obs_time, M)
    if proto == 2:
        n_obs = len(lt.obs_time)
        for m in range(M):
            reps["obs_begin"][m * n_rows:(m + 1) * n_rows] += m * n_obs
    steps = int(run["exp_type_fp"]["steps"])
    out = engine.run_replicas(reps, segs, steps, seed=99, obs_time=obs_time, trace=False, sync=True)
    out.raise_on_error()
    ref = mo.run(reps, segs, steps, seed=1234, obs_time=obs_time, parallel=True, trace=False)
    assert ref.rc == 0
    if proto == 1:
        g = out.final_n_e.reshape(M, n_rows).astype(float)
        o = ref.final_n_e.reshape(M, n_rows).astype(float)
    else:
        g = out.obs_n_e.reshape(M, -1).astype(float)
        o = ref.obs_n_e.reshape(M, -1).astype(float)
    for k in range(g.shape[1]):
        assert_means_agree(g[:, k], o[:, k], f"{exp}: column {k}")
    ge, oe = out.esteps.reshape(M, n_rows).sum(1).astype(float), ref.esteps.reshape(M, n_rows).sum(1).astype(float)
    assert_means_agree(ge, oe, f"{exp}: electron-steps per objective")


def test_simulate_api_in_philox_mode(gpu, tmp_path):
    """The public drop-in call with native random numbers: shapes, zero padding, CSV, reproducibility."""
    from mcluminescence_b200 import simulate as sim_mod
    from mcluminescence_b200.config import compose
    cfg = compose(overrides=["exp_type_fp.N_e=300", "exp_type_fp.holes=300", "exp_type_fp.steps=3000",
                             "exp_type_fp.T_rate=[5,20]", "exp_type_fp.duration=[120,40]", "exp_type_fp.sims=3",
                             "+tag=philox", "+seed=42"])
    sim_mod.PROJECT_ROOT = str(tmp_path)
    x_ax, lum, er, configs = sim_mod.simulate(cfg)
    assert x_ax.shape == lum.shape == er.shape == (3000, 3, 2) and len(configs) == 2
    for run in range(2):
        for j in range(3):
            n = int(np.count_nonzero(x_ax[:, j, run] > 0))
            assert 0 < n < 3000 and not x_ax[n:, j, run].any() and not lum[n:, j, run].any()
            assert np.all(np.diff(x_ax[:n, j, run]) >= 0) and x_ax[n - 1, j, run] >= configs[run].exp_type_fp.duration
            assert set(np.unique(lum[:n, j, run])) <= {0.0, 1.0}
            ne = np.rint(er[:n, j, run] * 300)
            assert np.array_equal(300 - np.cumsum(lum[:n, j, run]), ne)          # every event removes one electron
    import pandas as pd
    df = pd.read_csv(tmp_path / "results" / "simulations" / "exp_philox.csv")
    assert list(df.columns) == ["run", "sim", "step", "lum", "electron_ratio"] and len(df) == 3000 * 3 * 2
    again = sim_mod.simulate(cfg, write_csv=False)
    assert np.array_equal(again[0], x_ax) and np.array_equal(again[1], lum)       # same seed -> same run
    other = sim_mod.simulate(cfg, seed=43, write_csv=False)
    assert not np.array_equal(other[1], lum)


def test_tltrapsim_api_in_philox_mode(gpu, capsys):
    from mcluminescence_b200.config import compose, initialize_runs
    from mcluminescence_b200.tl_trap_lab import TLTrapSim
    run = initialize_runs(compose(overrides=helpers.LAB_OVERRIDES))[0]
    sim = TLTrapSim(run, seed=7)
    a = sim.TL_lab("CLBR_IRSL50_0.25KperGy")
    b = TLTrapSim(run, seed=7).TL_lab("CLBR_IRSL50_0.25KperGy")
    c = TLTrapSim(run, seed=8).ISO_lab("CLBR_IR50_ISO")
    assert a == b and 0.0 <= a < 1.0 and 0.0 <= c < 1.0 and sim.last_esteps > 0
    out = capsys.readouterr().out
    assert out.count("absError=") == 3 and "P_retrap=0.5" in out
    with pytest.raises(FileNotFoundError):
        TLTrapSim(run, seed=1).TL_lab("no_such_csv")
    with pytest.raises(TypeError):
        bad = initialize_runs(compose(overrides=["exp_type_fp=TLlab", "physics_fp=BG_basic"]))[0]
        TLTrapSim(bad)


@pytest.mark.parametrize("T_c,A_opt", [(50.0, 100.0), (250.0, 0.0)])
def test_decay_follows_the_nearest_neighbour_integral(gpu, T_c, A_opt):
    """Analytic pin (no oracle involved): with many more holes than electrons the kernel's decay must follow the
    nearest-neighbour integral of the localized-transition model for `k0 = A_opt + b e^{-E_loc/kT}` -- optical and
    thermal prefactor alike (same check as tests/test_oracle_analytic.py, same tolerance)."""
    from mcluminescence_b200 import engine
    from tests import test_oracle_analytic as ana
    cfg_overrides = [f"exp_type_fp.N_e={ana.N_E}", f"exp_type_fp.holes={ana.HOLES}", "exp_type_fp.e_ratio_start=1.0",
                     f"exp_type_fp.T_start=[{T_c}]", "exp_type_fp.T_rate=[0]", "exp_type_fp.duration=[1]",
                     "exp_type_fp.steps=400", "exp_type_fp.sims=1"]
    reps, segs, steps = ensemble_tables(cfg_overrides, ana.R)
    b, E, kb, rho_p = 1e12, 1.2, 8.617343e-05, 3e-4            # shipped basicTL12 / TL12 values (asserted below)
    from mcluminescence_b200.config import compose
    cfg = compose(overrides=cfg_overrides)
    assert (float(cfg["physics_fp"]["b"]), float(cfg["physics_fp"]["E_loc_1"]), float(cfg["physics_fp"]["k_b"]),
            float(cfg["exp_type_fp"]["rho_prime"])) == (b, E, kb, rho_p)
    k0 = A_opt + b * np.exp(-E / (kb * (T_c + 273.15)))
    grid = np.array([1e3, 1e4, 1e5, 1e6, 3e6, 1e7]) / k0
    segs["A_opt"] = A_opt
    segs["duration"] = float(grid[-1])
    out = engine.run_replicas(reps, segs, steps, seed=int(5 + T_c), sync=True)
    out.raise_on_error()
    frac = np.zeros((ana.R, len(grid)))
    for r in range(ana.R):
        n = int(out.steps_used[r])
        t, ne = out.t[r, :n], out.n_e[r, :n]
        idx = np.searchsorted(t, grid, side="right") - 1
        frac[r] = np.where(idx >= 0, ne[np.maximum(idx, 0)], ana.N_E) / ana.N_E
    got, se = frac.mean(axis=0), frac.std(axis=0, ddof=1) / np.sqrt(ana.R)
    want = ana.survival(k0, grid, rho_p)
    assert np.all(np.abs(got - want) <= 4.0 * se + 0.004), (got, want, se)
    # the specialised step loop (no trace): final occupancy only
    fin = engine.run_replicas(reps, segs, steps, seed=int(5 + T_c), trace=False, sync=True)
    fin.raise_on_error()
    assert np.array_equal(fin.final_n_e, out.final_n_e)


@pytest.mark.parametrize("retrap", [3.5 / 512.0, 0.1, 0.5, 0.77])
def test_channel_selector_matches_oracle_across_retrap(gpu, retrap):
    """The channel selector U < Retrap (engine.py:72) comes from the 9 spare low bits of the word that carries the
    exponential draw, ties settled by one more word: exact for every Retrap.  Channel 2 (E_loc_2 = 1.0 eV) tunnels
    ~80x faster than channel 1 at 250 degC, so the decay rate is nearly proportional to P(channel 2): 3.5/512 is the
    case in which one seventh of that probability comes from the tie-break word, 0.5 the one without any tie-break."""
    from mcluminescence_b200 import engine
    from oracle import mcl_oracle as mo
    overrides = list(CASES["iso_two_channel"][0])
    overrides[-1] = f"physics_fp.Retrap={retrap!r}"
    R = 512
    reps, segs, steps = ensemble_tables(overrides, R)
    out = engine.run_replicas(reps, segs, steps, seed=int(retrap * 1e6) + 3, sync=True)
    out.raise_on_error()
    ref = mo.run(reps, segs, steps, seed=4321, parallel=True)
    assert ref.rc == 0
    grid = np.geomspace(1, 900, 10)
    g = stats(out.event, out.n_e, out.t, out.steps_used, grid, int(reps['n_e0'][0]))
    o = stats(ref.event, ref.n_e, ref.t, ref.steps_used, grid, int(reps['n_e0'][0]))
    assert_means_agree(g[0], o[0], f"Retrap={retrap}: events per replica")
    for k in range(len(grid)):
        assert_means_agree(g[2][:, k], o[2][:, k], f"Retrap={retrap}: n_e(t={grid[k]:.3g})")
    assert g[0].mean() > 3.0                      # the case is informative: several events per replica
