"""Host-side logic that needs no GPU: config surface, sweep expansion, replica tables, replay cursor."""
import numpy as np
import pytest

from mcluminescence_b200.config import compose, initialize_runs, physics_record
from mcluminescence_b200.engine import ReplayStream
from mcluminescence_b200.replicas import LAB_CSV, LabTable, simulate_tables
from tests import helpers


def test_default_composition_matches_reference_surface():
    cfg = compose()
    assert set(cfg.keys()) == {"exp_type_fp", "physics_fp"}
    mc, ph = cfg.exp_type_fp, cfg.physics_fp
    assert mc.T_rate == [0.1, 1, 5, 20] and mc.duration == [8000, 800, 160, 40]
    assert mc.N_e == 2000 and mc.holes == 2000 and mc.steps == 20000 and mc.sims == 2
    # exponent forms without a dot must load as floats (OmegaConf's resolver), not strings
    assert isinstance(mc.rho_prime, float) and mc.rho_prime == 3e-4
    assert isinstance(ph.b, float) and ph.b == 1e12 and ph.s == 1e12 and ph.E_cb == 1000
    assert ph.k_b == 8.617343e-5


def test_overrides_and_group_swaps():
    cfg = compose(overrides=["exp_type_fp=TLlab", "physics_fp=lab_TL", "physics_fp.D=0.5",
                             "exp_type_fp.T_rate=[20]", "+tag=abc", "task=train"])
    assert cfg.exp_type_fp.N_e == 100 and cfg.physics_fp.D == 0.5 and cfg.physics_fp.E_cb == 1.8
    assert cfg.exp_type_fp.T_rate == [20] and cfg.get("tag", "") == "abc" and cfg.task == "train"
    assert cfg.get("missing", 7) == 7
    with pytest.raises(FileNotFoundError):
        compose(overrides=["physics_fp=nope"])


def test_initialize_runs_zip_cycles_and_rejects_ragged():
    runs = initialize_runs(compose())
    assert len(runs) == 4
    assert [runs[i].exp_type_fp.T_rate for i in range(4)] == [0.1, 1, 5, 20]
    assert [runs[i].exp_type_fp.duration for i in range(4)] == [8000, 800, 160, 40]
    assert all(runs[i].exp_type_fp.T_start == 0 and runs[i].exp_type_fp.boundary_factor == 1.2 for i in range(4))
    with pytest.raises(ValueError, match="not divisible"):
        initialize_runs(compose(overrides=["exp_type_fp.duration=[1,2,3]"]))
    assert len(initialize_runs(compose(overrides=["exp_type_fp=TLlab", "physics_fp=lab_TL"]))) == 1


def test_physics_record_rejects_unknown_and_missing_keys_like_the_dataclass():
    with pytest.raises(TypeError, match="unexpected keyword argument 'E'"):
        physics_record(compose(overrides=["physics_fp=BG_basic"]).physics_fp)
    with pytest.raises(TypeError, match="missing"):
        physics_record({"alpha": 1.0})
    with pytest.raises(TypeError):          # simulate()'s dose_rate key can never reach Physics
        physics_record(compose(overrides=["+physics_fp.dose_rate=0.1"]).physics_fp)


def test_replica_tables_follow_reference_truncations():
    runs = initialize_runs(compose())
    reps, segs = simulate_tables(runs, 2)
    assert len(reps) == 8 and len(segs) == 4
    assert set(reps["n_h0"]) == {3455}              # int(2000 * 1.2**3), not 3456
    assert set(reps["n_e0"]) == {2000} and set(reps["N_e"]) == {2000}
    np.testing.assert_allclose(reps["side"], 1.8962e-7, rtol=1e-4)
    assert list(reps["seg_begin"]) == [0, 0, 1, 1, 2, 2, 3, 3]
    assert list(segs["dt_cap"]) == [1 / 0.1, 1.0, 1 / 5, 1 / 20]
    r0 = initialize_runs(compose(overrides=["exp_type_fp.T_rate=[0]", "exp_type_fp.duration=[5]"]))
    assert simulate_tables(r0, 1)[1]["dt_cap"][0] == 1e20


def test_lab_tables():
    run = initialize_runs(compose(overrides=helpers.LAB_OVERRIDES))[0]
    tl = LabTable(*LAB_CSV["tl_clbr"], helpers.DATA_ROOT)
    reps, segs = tl.tables(run)
    assert len(reps) == 11 and set(reps["n_e0"]) == {0} and set(reps["n_h0"]) == {172}
    assert np.all(segs["dose_rate"] == 0.092) and np.all(segs["T_rate"] < 0)
    iso = LabTable(*LAB_CSV["iso"], helpers.DATA_ROOT)
    reps, segs = iso.tables(run)
    assert len(reps) == 8 and iso.obs_begin[-1] == 91 and len(iso.target) == 91
    assert segs["dose_rate"][0] == 0.092 and np.all(segs["dose_rate"][1:] == 0)
    assert list(reps["n_e0"]) == [0, 95, 94, 94, 94, 91, 53, 4]
    er, mse = tl.mse(100, np.full(11, 50))
    assert mse == float(np.mean([(0.5 - f) ** 2 for f in tl.target]))


def test_replay_stream_is_numpys_legacy_stream_with_a_cursor():
    ref = np.random.RandomState(3).random_sample(3_100_000)
    st = ReplayStream(3)
    st.advance(516)
    assert np.array_equal(st.window(100), ref[516:616])
    st.advance(70)
    assert np.array_equal(st.window(2_000_000), ref[586:2_000_586])
    st.advance(3_000_000)
    assert np.array_equal(st.window(10), ref[3_000_586:3_000_596])


def test_glow_curve_smoothing_matches_reference_definition():
    from mcluminescence_b200.postprocess import decay_curve, glow_curve, running_mean
    rs = np.random.RandomState(0)
    hist = rs.poisson(20, size=300)
    g = glow_curve(hist, n_replicas=10, bin_width=1.0, win_deg=50.0)
    assert g.shape == (251,)
    assert np.allclose(g[0], hist[:50].mean() / 10)
    assert np.allclose(running_mean(np.arange(10.0), 5), np.arange(2.0, 8.0))
    occ = np.array([40, 20]); occ2 = np.array([40 * 40 // 4 * 1 + 0, 120])   # 4 replicas
    m, s = decay_curve(occ, np.array([400, 120]), 4, 10.0)
    assert np.allclose(m, [1.0, 0.5]) and np.allclose(s, [0.0, np.sqrt(120 / 4 - 25) / 10])
