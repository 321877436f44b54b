"""Analytic pins of the CPU oracle for the one piece of the path the reference source does not contain: the
optical-stimulation prefactor `k_tun = (A_opt + b e^{-E_loc/kT}) e^{-alpha r}` ("parity unpinned").

With many more holes than electrons every electron keeps its own nearest hole, so the decay is a superposition of
independent first-order decays over the nearest-neighbour distance distribution of a random hole cloud,

    n(t) / n0 = Integral 3 rho' r'^2 exp(-rho' r'^3) exp(-k0 t e^{-r'}) dr',      r' = alpha r,

the localized-transition model (Huntley 2006; Jain, Guralnik & Andersen 2012) that MCLuminescence simulates; rho' is the
`rho_prime` of the config (`tl_trap_lab.py:33`).  The same integral pins the thermal tunnelling prefactor with
`k0 = b e^{-E_loc/kT}` (A_opt = 0), i.e. the reference's own rate law, as a cross-check of the method."""
import numpy as np
import pytest

from mcluminescence_b200.config import compose, initialize_runs
from mcluminescence_b200.replicas import simulate_tables
from oracle import mcl_oracle as mo

N_E, HOLES, R = 100, 20000, 96


def survival(k0, t, rho_p):
    r = np.linspace(0.0, 80.0, 400001)
    pdf = 3.0 * rho_p * r ** 2 * np.exp(-rho_p * r ** 3)
    return np.array([np.trapezoid(pdf * np.exp(-k0 * tt * np.exp(-r)), r) for tt in t])


def run_case(T_c, A_opt, duration, seed):
    cfg = compose(overrides=[f"exp_type_fp.N_e={N_E}", f"exp_type_fp.holes={HOLES}", "exp_type_fp.e_ratio_start=1.0",
                             f"exp_type_fp.T_start=[{T_c}]", "exp_type_fp.T_rate=[0]", f"exp_type_fp.duration=[{duration}]",
                             "exp_type_fp.steps=400", "exp_type_fp.sims=1"])
    run = initialize_runs(cfg)
    reps1, segs = simulate_tables(run, 1)
    segs["A_opt"] = A_opt
    reps = np.repeat(reps1, R)
    res = mo.run(reps, segs, 400, seed=seed, parallel=True)
    assert res.rc == 0
    return run[0], res


@pytest.mark.parametrize("T_c,A_opt", [(50.0, 100.0), (50.0, 3.0), (250.0, 0.0)])
def test_decay_follows_the_nearest_neighbour_integral(T_c, A_opt):
    cfgrun, _ = run_case(T_c, A_opt, 1.0, seed=1)          # only to read the physics record
    ph, mc = cfgrun["physics_fp"], cfgrun["exp_type_fp"]
    k0 = A_opt + float(ph["b"]) * np.exp(-float(ph["E_loc_1"]) / (float(ph["k_b"]) * (T_c + 273.15)))
    rho_p = float(mc["rho_prime"])
    # times that span survival ~0.87 .. ~0.25
    grid = np.array([1e3, 1e4, 1e5, 1e6, 3e6, 1e7]) / k0
    _, res = run_case(T_c, A_opt, float(grid[-1]), seed=int(7 + T_c + A_opt))
    frac = np.zeros((R, len(grid)))
    for r in range(R):
        n = int(res.steps_used[r])
        t, ne = res.t[r, :n], res.n_e[r, :n]
        idx = np.searchsorted(t, grid, side="right") - 1
        frac[r] = np.where(idx >= 0, ne[np.maximum(idx, 0)], N_E) / N_E
    got, se = frac.mean(axis=0), frac.std(axis=0, ddof=1) / np.sqrt(R)
    want = survival(k0, grid, rho_p)
    assert 0.15 < want[-1] < 0.45 and want[0] > 0.8 and np.all(np.diff(want) < 0)
    # hole depletion (an electron loses its nearest hole to another one: 100 electrons among 3.5e4 holes) and the finite
    # box bias the simulated survival upwards by a few tenths of a percent at most
    assert np.all(np.abs(got - want) <= 4.0 * se + 0.004), (got, want, se)
