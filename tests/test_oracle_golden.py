"""Pin the CPU oracle (oracle/mcl_oracle.c) against golden vectors from the UNMODIFIED reference.

The vectors under tests/golden/ were produced by oracle/ref_harness/gen_golden.py, which runs the
reference's own simulate() / optimizer.objective() with np.random.seed(...).  Integer traces and
the structural (electron index, hole index) log must match bit-for-bit; event times to 1e-12
relative (NumPy's AVX512 exp differs from libm by 1 ulp on ~5 % of arguments).
"""
import numpy as np
import pytest

from oracle import mcl_oracle as mo
from tests import helpers

T_RTOL = 1e-12

SIM_FAST = ["sim_kat1", "sim_kat2", "sim_sweep2x2", "sim_partial", "sim_empty", "sim_labphys",
            "sim_zero_duration"]


def check_sim_case(golden, name):
    meta, arrs = golden.meta(name), golden.arrays(name)
    reps, segs, steps, sims, n_runs = helpers.sim_tables_for(meta)
    res = mo.run(reps, segs, steps, seed=meta["seed"])
    assert res.rc == 0
    assert list(res.steps_used) == list(arrs["steps_used"])
    assert int(res.consumed.sum()) == meta["n_uniforms"]
    for r, (ev, ne, tt) in enumerate(helpers.split_traces(arrs)):
        n = len(ev)
        assert np.array_equal(res.event[r, :n], ev), f"{name} replica {r}: event trace"
        assert np.array_equal(res.n_e[r, :n], ne), f"{name} replica {r}: n_e trace"
        np.testing.assert_allclose(res.t[r, :n], tt, rtol=T_RTOL, atol=0)
    # structural log: seeds (kind 0) interleaved with fills (1) / recombinations (2)
    log = arrs["log"]
    seeds = log[log[:, 0] == 0]
    assert np.array_equal(seeds[:, 1], reps["n_e0"]) and np.array_equal(seeds[:, 2], reps["n_h0"])
    got = np.concatenate(helpers.log_from_result(res, range(len(reps))))
    assert np.array_equal(got, log[log[:, 0] > 0])


@pytest.mark.parametrize("name", SIM_FAST)
def test_simulate_cases(golden, name):
    check_sim_case(golden, name)


def test_simulate_kat0_default_config(golden):
    """SURVEY KAT-0: the shipped default config, 8 replicas on one stream, 19 014 662 electron-steps."""
    if "sim_kat0" not in golden.manifest:
        pytest.skip("kat0 golden not generated")
    check_sim_case(golden, "sim_kat0")
    meta = golden.meta("sim_kat0")
    reps, segs, steps, sims, n_runs = helpers.sim_tables_for(meta)
    res = mo.run(reps, segs, steps, seed=meta["seed"], trace=False)
    assert int(res.esteps.sum()) == 19014662
    assert [int(v) for v in res.steps_used] == [1807, 1833, 1710, 1705, 1645, 1636, 1582, 1569]


def run_lab_oracle(meta):
    lt, reps, segs, run = helpers.lab_setup(meta)
    rng = mo.Rng(meta["seed"])
    # run_one_sim first builds TLTrapSim(runs[0]) (optimizer.py:71): Box.seed(0, holes) draws
    # 3 * n_h0 uniforms that no lab row uses.
    rng.uniforms(3 * int(reps["n_h0"][0]))
    res = mo.run(reps, segs, int(run["exp_type_fp"]["steps"]), rng=rng, obs_time=lt.obs_time)
    return lt, reps, run, res, rng


@pytest.mark.parametrize("name", [
    "lab_default_tl_clbr", "lab_default_tl_fsm13", "lab_default_iso",
    "lab_kat3_tl_clbr", "lab_kat3_tl_fsm13", "lab_kat3_iso",
    "lab_sobol0_tl_clbr", "lab_sobol1_tl_clbr", "lab_sobol2_tl_clbr", "lab_sobol3_tl_clbr",
    "lab_sobol0_iso", "lab_sobol1_iso"])
def test_lab_cases(golden, name):
    meta, arrs = golden.meta(name), golden.arrays(name)
    lt, reps, run, res, rng = run_lab_oracle(meta)
    assert res.rc == 0
    assert rng.consumed == meta["n_uniforms"]
    log = arrs["log"]
    got = np.concatenate(helpers.log_from_result(res, range(len(reps))))
    assert np.array_equal(got, log[log[:, 0] > 0])
    _, mse = lt.mse(run["exp_type_fp"]["N_e"], res.final_n_e, res.obs_n_e)
    assert mse == meta["value"]          # same integers -> same float arithmetic -> same bits


def test_mt19937_matches_numpy_legacy_stream():
    for seed in (0, 1, 12345, 2**32 - 1):
        u = mo.Rng(seed).uniforms(2000)
        assert np.array_equal(u, np.random.RandomState(seed).random_sample(2000))


def test_external_stream_equals_seeded_stream(golden):
    meta = golden.meta("sim_kat2")
    reps, segs, steps, _, _ = helpers.sim_tables_for(meta)
    a = mo.run(reps, segs, steps, seed=meta["seed"])
    u = np.random.RandomState(meta["seed"]).random_sample(int(a.consumed.sum()))
    b = mo.run(reps, segs, steps, rng=mo.Rng(external=u))
    assert b.rc == 0 and np.array_equal(a.event, b.event) and np.array_equal(a.t, b.t)
    short = mo.run(reps, segs, steps, rng=mo.Rng(external=u[:-5]))
    assert short.rc == -3


def test_steps_overflow_is_reported(golden):
    meta = golden.meta("sim_kat1")
    reps, segs, steps, _, _ = helpers.sim_tables_for(meta)
    res = mo.run(reps, segs, 100, seed=meta["seed"])
    assert res.rc == -1 and res.status[0] == -1


def test_parallel_run_is_deterministic_per_replica(golden):
    meta = golden.meta("sim_sweep2x2")
    reps, segs, steps, _, _ = helpers.sim_tables_for(meta)
    a = mo.run(reps, segs, steps, seed=9, parallel=True, threads=4)
    b = mo.run(reps, segs, steps, seed=9, parallel=True, threads=1)
    assert a.rc == 0 and np.array_equal(a.event, b.event) and np.array_equal(a.n_e, b.n_e)
    # replica r of the parallel run is the sequential run of that replica alone on stream seed+r
    c = mo.run(reps[2:3], segs, steps, seed=9 + 2)
    assert np.array_equal(a.event[2], c.event[0])
