"""GPU parity AT THE BENCHMARKED SHAPES: the kernel instantiations and box sizes `bench.py` times, against the CPU
oracle (float64, MT19937) run on the same replica tables.

* C2 -- 10^4-electron boxes (17 279 holes), hold + optical readout, `philox_kernel<256,3,uint16_t,2>`, the specialised
  step loop with fused histograms (what the bench launches): per-leg event histograms and n(t) on the bench's own
  log-time axis, plus a Kolmogorov-Smirnov test on event times from a traced run of the same instantiation.
* C5 -- 2000-electron boxes on a 1 degC/s ramp, `philox_kernel<64,16,uint16_t,2>`: glow curve and n(T).
* C3 -- irradiation from empty traps (dose in the simulate protocol, fills, stale caches) followed by a TL ramp on the
  same box (reference simulate.py:42,58,70-73 with engine.py:133-152), per dose group.
* a conduction-band-dominated 10^4-electron box, whose decay is exp(-k_cb t) analytically: bounds the bias the 24-bit
  uniforms and the SFU logarithm can put on the minimum of 10^4 clocks.
* the small-waiting-time tail of the SFU exponential draw itself, against float64.

Tolerances: ensemble means within 4.5 combined standard errors (hundreds of correlated bins are checked per case;
north_star's "3 sigma of the replica spread" is far looser), KS p > 1e-3.
"""
import os

import numpy as np
import pytest

from tests import helpers

pytestmark = pytest.mark.gpu

Z = 4.5


@pytest.fixture(scope="module")
def gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("marked gpu but no CUDA device is visible")
    from mcluminescence_b200 import _native
    _native.load()
    return torch


class env:
    def __init__(self, **kv):
        self.kv = kv

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kv}
        os.environ.update({k: str(v) for k, v in self.kv.items()})

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def compare_with_oracle(wl, out, ref, R_g, what, coarse=40):
    """GPU fused histograms (sums over R_g replicas) vs the oracle's traces rebinned per replica."""
    hist, segs = wl["hist"], wl["segments"]
    n0 = int(wl["replicas"]["n_e0"][0])
    R_o = ref.event.shape[0]
    ev_o, occ_o, n_end_o, n_ev_o, times_o = helpers.leg_histograms(ref.event, ref.n_e, ref.t, ref.steps_used, n0, segs, hist)
    g_ev, g_occ, g_occ2 = out.hist_events.astype(float), out.hist_occ.astype(float), out.hist_occ_sq.astype(float)
    worst = 0.0
    for sg in range(len(segs)):
        edges = helpers.axis_edges_in_time(hist, segs[sg])
        full = edges[:-1] <= float(segs["duration"][sg])            # edges every replica crosses
        # ---- n(t): mean occupancy at the bin edges
        m_g = g_occ[sg, full] / R_g
        v_g = np.maximum(g_occ2[sg, full] / R_g - m_g * m_g, 0.0)
        o = occ_o[:, sg, :][:, full]
        assert not np.isnan(o).any(), f"{what}: an oracle replica ended leg {sg} before an edge inside the leg"
        m_o, v_o = o.mean(0), o.var(0, ddof=1)
        se = np.sqrt(v_g / R_g + v_o / R_o)
        z = np.abs(m_g - m_o) / (se + 1e-9)
        bad = np.abs(m_g - m_o) > Z * se + 0.5
        assert not bad.any(), (f"{what}: n(t) leg {sg}: {int(bad.sum())} of {bad.size} edges off, worst {z[bad].max():.1f} sigma "
                               f"(GPU {m_g[bad][0]:.2f} vs oracle {m_o[bad][0]:.2f})")
        worst = max(worst, float(np.max(np.where(se > 0.05, z, 0.0))))
        # ---- L(t): events per coarse bin and replica
        nb = hist.n_bins // coarse * coarse
        e_g = g_ev[sg, :nb].reshape(-1, coarse).sum(1) / R_g
        e_o = ev_o[:, sg, :nb].reshape(R_o, -1, coarse).sum(2)
        m_o, v_o = e_o.mean(0), e_o.var(0, ddof=1)
        se = np.sqrt(np.maximum(v_o, m_o + 1e-9) * (1.0 / R_g + 1.0 / R_o))      # at least Poisson
        bad = np.abs(e_g - m_o) > Z * se + 0.05
        assert not bad.any(), (f"{what}: L(t) leg {sg}: coarse bins {np.nonzero(bad)[0]} off: GPU {e_g[bad]} vs oracle {m_o[bad]} "
                               f"(se {se[bad]})")
        assert e_g.sum() > 0 and m_o.sum() > 0, f"{what}: leg {sg} produced no luminescence"
    return times_o, n_end_o, n_ev_o, worst


def test_c2_bench_instantiation_matches_oracle(gpu):
    """BASELINE config 2 exactly as bench.py launches it (more than 296 replicas => 256-thread CTAs, three per SM,
    specialised step loop, fused histograms), against 40 oracle replicas of the same 10^4-electron table."""
    from mcluminescence_b200 import engine, workloads
    from oracle import mcl_oracle as mo
    from scipy.stats import ks_2samp
    R_g, R_o, R_t = 320, 40, 48
    wl = workloads.c2(n_replicas=R_g)
    assert int(wl["replicas"]["n_e0"][0]) == 10_000 and int(wl["replicas"]["n_h0"][0]) == 17_279
    out = engine.run_replicas(wl["replicas"], wl["segments"], wl["max_steps"], seed=1201, hist=wl["hist"], trace=False, sync=True)
    out.raise_on_error()
    ref = mo.run(wl["replicas"][:R_o], wl["segments"], wl["max_steps"], seed=88, parallel=True)
    assert ref.rc == 0
    times_o, n_end_o, n_ev_o, worst = compare_with_oracle(wl, out, ref, R_g, "C2")
    # final occupancy and electron-steps per replica
    fin_g, fin_o = out.final_n_e.astype(float), ref.final_n_e.astype(float)
    se = np.sqrt(fin_g.var(ddof=1) / R_g + fin_o.var(ddof=1) / R_o)
    assert abs(fin_g.mean() - fin_o.mean()) <= Z * se, (fin_g.mean(), fin_o.mean(), se)
    es_g, es_o = out.esteps.astype(float), ref.esteps.astype(float)
    se = np.sqrt(es_g.var(ddof=1) / R_g + es_o.var(ddof=1) / R_o)
    assert abs(es_g.mean() - es_o.mean()) <= Z * se, (es_g.mean(), es_o.mean(), se)
    # the same instantiation with per-step records (general step loop): identical replicas, and event times for KS
    with env(MCL_PHILOX_NT=256):
        tr = engine.run_replicas(wl["replicas"][:R_t], wl["segments"], wl["max_steps"], seed=1201, trace=True, sync=True)
    tr.raise_on_error()
    assert np.array_equal(tr.final_n_e, out.final_n_e[:R_t]) and np.array_equal(tr.esteps, out.esteps[:R_t])
    _, _, _, _, times_g = helpers.leg_histograms(tr.event, tr.n_e, tr.t, tr.steps_used, 10_000, wl["segments"], wl["hist"])
    for sg in range(2):
        # events of one replica are correlated: thin both pools so that the KS p-value stays meaningful
        a, b = np.sort(times_g[sg])[::16], np.sort(times_o[sg])[::16]
        ks = ks_2samp(a, b)
        assert ks.pvalue > 1e-3, f"C2 leg {sg}: KS on event times p={ks.pvalue:.2e} D={ks.statistic:.4f}"
    print(f"C2 bench shape: worst n(t) deviation {worst:.2f} sigma; final n_e GPU {fin_g.mean():.1f} oracle {fin_o.mean():.1f}")


def test_c5_bench_instantiation_matches_oracle(gpu):
    """BASELINE config 5's replica shape (2000 electrons, 1 degC/s ramp) on the 64-thread instantiation the bench
    launches, fused temperature histograms, against 64 oracle replicas."""
    from mcluminescence_b200 import engine, workloads
    from oracle import mcl_oracle as mo
    from scipy.stats import ks_2samp
    R_g, R_o, R_t = 640, 64, 96
    wl = workloads.c5(n_replicas=R_g)
    out = engine.run_replicas(wl["replicas"], wl["segments"], wl["max_steps"], seed=1305, hist=wl["hist"], trace=False, sync=True)
    out.raise_on_error()
    ref = mo.run(wl["replicas"][:R_o], wl["segments"], wl["max_steps"], seed=99, parallel=True)
    assert ref.rc == 0
    times_o, _, _, worst = compare_with_oracle(wl, out, ref, R_g, "C5", coarse=20)
    fin_g, fin_o = out.final_n_e.astype(float), ref.final_n_e.astype(float)
    se = np.sqrt(fin_g.var(ddof=1) / R_g + fin_o.var(ddof=1) / R_o)
    assert abs(fin_g.mean() - fin_o.mean()) <= Z * se, (fin_g.mean(), fin_o.mean(), se)
    st_g, st_o = out.steps_used.astype(float), ref.steps_used.astype(float)
    se = np.sqrt(st_g.var(ddof=1) / R_g + st_o.var(ddof=1) / R_o)
    assert abs(st_g.mean() - st_o.mean()) <= Z * se, (st_g.mean(), st_o.mean(), se)
    with env(MCL_PHILOX_NT=64):
        tr = engine.run_replicas(wl["replicas"][:R_t], wl["segments"], wl["max_steps"], seed=1305, trace=True, sync=True)
    tr.raise_on_error()
    assert np.array_equal(tr.final_n_e, out.final_n_e[:R_t]) and np.array_equal(tr.steps_used, out.steps_used[:R_t])
    _, _, _, _, times_g = helpers.leg_histograms(tr.event, tr.n_e, tr.t, tr.steps_used, 2000, wl["segments"], wl["hist"])
    ks = ks_2samp(np.sort(times_g[0])[::8], np.sort(times_o[0])[::8])
    assert ks.pvalue > 1e-3, f"C5: KS on event times p={ks.pvalue:.2e} D={ks.statistic:.4f}"
    print(f"C5 bench shape: worst n(T) deviation {worst:.2f} sigma")


def test_pipelined_two_channel_and_ramp_match_oracle(gpu):
    """The pipelined step loop against the oracle DIRECTLY (its equality with the loop in order is tested in
    test_gpu_ensemble.py) on the two forms the bench shapes do not reach: two distinct tunnelling channels (hold + optical
    readout, 4000-electron boxes => 128-thread CTAs, one Philox call per slot pair, per-electron selector), and a ramp
    (4000 electrons, 2 degC/s: the sweep team's clocks carry no temperature, the decision warp adds the prefactor)."""
    from mcluminescence_b200 import engine, workloads
    from oracle import mcl_oracle as mo
    R_g, R_o = 192, 32
    two = workloads.c2(n_replicas=R_g, n_e=4000, n_bins=200, physics_overrides=["physics_fp.E_loc_2=1.0", "physics_fp.Retrap=0.3"])
    ramp = workloads.c5(n_replicas=R_g, n_bins=200)
    for k in ("N_e", "n_e0"):
        ramp["replicas"][k] = 4000
    ramp["replicas"]["n_h0"] = int(4000 * 1.2 ** 3)
    ramp["replicas"]["side"] = ramp["replicas"]["side"] * (4000 / 2000) ** (1.0 / 3.0)          # same hole density as C5
    ramp["segments"]["T_rate"] = 2.0
    ramp["segments"]["duration"] = 400.0
    ramp["segments"]["dt_cap"] = 0.5
    for what, wl, coarse in (("two channels, pipelined", two, 20), ("ramp, pipelined", ramp, 10)):
        with env(MCL_PHILOX_NT=128):
            out = engine.run_replicas(wl["replicas"], wl["segments"], wl["max_steps"], seed=1777, hist=wl["hist"], trace=False, sync=True)
            out.raise_on_error()
            with env(MCL_PHILOX_PIPE=0):
                inorder = engine.run_replicas(wl["replicas"][:32], wl["segments"], wl["max_steps"], seed=1777, hist=wl["hist"], trace=False, sync=True)
        assert np.array_equal(inorder.steps_used, out.steps_used[:32]) and np.array_equal(inorder.esteps, out.esteps[:32]), what
        ref = mo.run(wl["replicas"][:R_o], wl["segments"], wl["max_steps"], seed=91, parallel=True)
        assert ref.rc == 0
        _, _, _, worst = compare_with_oracle(wl, out, ref, R_g, what, coarse=coarse)
        fin_g, fin_o = out.final_n_e.astype(float), ref.final_n_e.astype(float)
        se = np.sqrt(fin_g.var(ddof=1) / R_g + fin_o.var(ddof=1) / R_o)
        assert abs(fin_g.mean() - fin_o.mean()) <= Z * se + 0.5, (what, fin_g.mean(), fin_o.mean(), se)
        es_g, es_o = out.esteps.astype(float), ref.esteps.astype(float)
        se = np.sqrt(es_g.var(ddof=1) / R_g + es_o.var(ddof=1) / R_o)
        assert abs(es_g.mean() - es_o.mean()) <= Z * se, (what, es_g.mean(), es_o.mean(), se)
        assert int(out.steps_used.min()) > 500, what              # long enough to have run pipelined
        print(f"{what}: worst n(t) deviation {worst:.2f} sigma; final n_e GPU {fin_g.mean():.1f} oracle {fin_o.mean():.1f}")


def test_c3_dose_then_tl_matches_oracle(gpu):
    """BASELINE config 3: irradiation from EMPTY traps in the simulate protocol (`dose_rate` != 0: fills, stale
    incremental cache, the h+1 re-scan -- simulate.py:42,58,70-73, engine.py:133-175), then a TL ramp on the same box
    in fill mode.  Per dose group: fill after the irradiation, TL events, the glow curve, and the fused histogram."""
    from mcluminescence_b200 import engine, workloads
    from oracle import mcl_oracle as mo
    per_g, per_o = 24, 12
    wl = workloads.c3(replicas_per_dose=per_g)
    wl_o = workloads.c3(replicas_per_dose=per_o)
    segs = wl["segments"]
    out = engine.run_replicas(wl["replicas"], segs, wl["max_steps"], seed=1407, hist=wl["hist"], hist_group=wl["hist_group"],
                              trace=True, sync=True)
    out.raise_on_error()
    ref = mo.run(wl_o["replicas"], segs, wl["max_steps"], seed=123, parallel=True)
    assert ref.rc == 0

    def per_replica(res, reps):
        R = len(reps)
        fill_end, tl_events, glow, in_axis, irr_events = np.zeros(R), np.zeros(R), np.zeros((R, 16)), np.zeros(R), np.zeros(R)
        for r in range(R):
            sb = int(reps["seg_begin"][r])
            n = int(res.steps_used[r])
            tt, ee, nn = res.t[r, :n], res.event[r, :n], res.n_e[r, :n]
            (a0, a1, _), (b0, b1, t_off) = helpers.split_legs(tt, segs["duration"][sb:sb + 2])
            fill_end[r] = nn[a1 - 1]
            hit = ee[b0:b1] > 0
            tl_events[r] = hit.sum()
            T = (tt[b0:b1] - t_off)[hit] * 1.0                       # 0 degC start, 1 degC/s
            glow[r] = np.histogram(T, bins=np.linspace(0.0, 800.0, 17))[0]
            in_axis[r] = np.count_nonzero(T < 800.0)
            irr_events[r] = np.count_nonzero(ee[a0:a1] > 0)
        return fill_end, tl_events, glow, in_axis, irr_events

    fg, tg, gg, ag, ig = per_replica(out, wl["replicas"])
    fo, to, go, _, _ = per_replica(ref, wl_o["replicas"])
    assert fg.max() > 1500 and fg.min() > 0                              # the box really fills up (N_e = 2000)
    for d in range(10):
        sg_, so_ = slice(d * per_g, (d + 1) * per_g), slice(d * per_o, (d + 1) * per_o)
        for name, a, b in (("fill after irradiation", fg[sg_], fo[so_]), ("TL events", tg[sg_], to[so_])):
            se = np.sqrt(a.var(ddof=1) / per_g + b.var(ddof=1) / per_o)
            assert abs(a.mean() - b.mean()) <= Z * se + 0.5, f"C3 dose {d}: {name}: GPU {a.mean():.1f} vs oracle {b.mean():.1f} (se {se:.2f})"
        ga, gb = gg[sg_], go[so_]
        se = np.sqrt(np.maximum(ga.var(0, ddof=1), ga.mean(0) + 1e-9) / per_g + np.maximum(gb.var(0, ddof=1), gb.mean(0) + 1e-9) / per_o)
        bad = np.abs(ga.mean(0) - gb.mean(0)) > Z * se + 0.5
        assert not bad.any(), f"C3 dose {d}: glow-curve bins {np.nonzero(bad)[0]}: GPU {ga.mean(0)[bad]} vs oracle {gb.mean(0)[bad]}"
        # fused histogram rows of this dose group: leg 0 -> row 2d (isothermal at 15 degC: every event in bin 15),
        # leg 1 -> row 2d + 1 (every TL event below 800 degC)
        assert int(out.hist_events[2 * d + 1].sum()) == int(ag[sg_].sum())
        assert int(out.hist_events[2 * d].sum()) == int(out.hist_events[2 * d][15]) == int(ig[sg_].sum())
    # dose response: more irradiation, more trapped charge
    means = fg.reshape(10, per_g).mean(1)
    assert np.all(np.diff(means) > 0), means


def test_cb_dominated_large_box_decays_exponentially(gpu):
    """10^4 electrons whose rates are all (nearly) the conduction-band rate k_cb: E[n(t)] = n0 exp(-k_cb t) exactly,
    and every step is the minimum of ~10^4 EQUAL clocks -- the case in which the 24-bit uniforms and the absolute error
    of `lg2.approx` near u -> 1 matter most (E_min ~ 1/n).  The fitted rate must agree with k_cb to 0.3 %."""
    from mcluminescence_b200 import engine, workloads
    from mcluminescence_b200.engine import AXIS_TIME_LIN, HistSpec
    R = 600
    wl = workloads.c2(n_replicas=R, physics_overrides=["physics_fp.E_cb=1.0"])
    rp = wl["replicas"][0]
    T = 250.0 + 273.15
    k_cb = float(rp["s"]) * np.exp(-float(rp["E_cb"]) / (float(rp["k_b"]) * T))
    k_tun_max = float(rp["b"]) * np.exp(-float(rp["E_loc_1"]) / (float(rp["k_b"]) * T))
    assert k_tun_max < 0.02 * k_cb                                  # tunnelling is a < 1e-4 correction on average
    segs = wl["segments"][:1].copy()
    segs["duration"] = 2.0 / k_cb
    reps = wl["replicas"].copy()
    reps["seg_count"] = 1
    hist = HistSpec(axis=AXIS_TIME_LIN, n_bins=40, lo=0.0, hi=2.0 / k_cb, n_groups=1)
    out = engine.run_replicas(reps, segs, wl["max_steps"], seed=1509, hist=hist, trace=False, sync=True)
    out.raise_on_error()
    t_edges = np.arange(40) * (2.0 / k_cb / 40)
    frac = out.hist_occ[0] / (R * 1e4)
    want = np.exp(-k_cb * t_edges)
    se = np.sqrt(want * (1 - want) / (R * 1e4))
    # fitted rate from the late edges (log-linear, weighted)
    sel = slice(8, 40)
    k_fit = -np.sum(np.log(frac[sel]) * t_edges[sel]) / np.sum(t_edges[sel] ** 2)
    bias = k_fit / k_cb - 1.0
    print(f"CB-dominated 10^4-electron box: fitted rate / k_cb - 1 = {bias:+.2e}")
    assert abs(bias) < 3e-3, bias
    assert np.all(np.abs(frac - want) <= Z * se + 3e-3 * want * k_cb * t_edges + 1e-6), (frac - want, se)


def test_small_waiting_time_tail_of_the_sfu_draw(gpu):
    """-lg2.approx(u) for u -> 1 (PTX guarantees only an ABSOLUTE error of 2^-22 there, the size of the smallest
    draws themselves): every draw must stay positive and finite, the tail E < 1e-5 must follow float64 closely, and
    the dose-free filling clock's lower bound must clear the kernel's 5e12 s skip threshold."""
    import ctypes as C
    from mcluminescence_b200 import _native
    L = _native.load()
    top = np.arange(2 ** 23 - 4096, 2 ** 23, dtype=np.uint64)                   # the 4096 largest mantissas: u closest to 1
    rng = np.random.default_rng(5)
    mant = np.concatenate([top, rng.integers(2 ** 23 - 2 ** 17, 2 ** 23, 200_000, dtype=np.uint64),
                           rng.integers(0, 2 ** 23, 200_000, dtype=np.uint64), np.arange(0, 64, dtype=np.uint64)])
    words = (mant << np.uint64(9)).astype(np.uint32)
    a = np.zeros(words.size, np.float32)
    b = np.zeros(words.size, np.float32)
    _native.check(L.mcl_debug_exp_draws(words.ctypes.data, int(words.size), a.ctypes.data, b.ctypes.data), "mcl_debug_exp_draws")
    u = (2.0 * mant.astype(np.float64) + 1.0) * 2.0 ** -24                       # u01() exactly
    exact = -np.log2(u)
    assert np.all(np.isfinite(a)) and np.all(a > 0), "a draw came out zero, negative or non-finite"
    assert np.all(np.isfinite(b))
    err = a.astype(np.float64) - exact
    rel = err / exact
    near1 = u > 0.5
    tail = exact < 1e-5 / np.log(2.0)
    print(f"SFU draw: abs. error for u in (0.5, 1): mean {err[near1].mean():+.2e} max {np.abs(err[near1]).max():.2e} (2^-22 = {2.0**-22:.2e}); "
          f"tail (E<1e-5) rel. error mean {rel[tail].mean():+.2e}; smallest draw {a.min():.3e} (exact {exact.min():.3e})")
    # what PTX promises near u -> 1 is an ABSOLUTE error of 2^-22; measured on B200: a near-constant -7e-8, i.e. draws of
    # size E are short by 7e-8 / E -- 5e-4 of the minimum of 10^4 equal clocks, the worst case the path has
    # (test_cb_dominated_large_box_decays_exponentially bounds its effect on the decay rate)
    assert np.abs(err[near1]).max() < 2.0 ** -22
    assert np.abs(err[near1 & (exact > 1e-4)] / exact[near1 & (exact > 1e-4)]).max() < 1.5e-3
    assert np.abs(rel[~near1]).max() < 1e-5
    # outer logarithm (what enters the argmin): absolute error in log2 units
    assert np.abs(b.astype(np.float64) - np.log2(a.astype(np.float64))).max() < 2e-5
    # dose-free filling clock exponential(1e20 s): the kernel clamps its draw at 5e-8 (below the exact minimum 8.6e-8,
    # above the SFU's worst) and skips the clock while something else happens before 3e12 s
    assert exact.min() > 5.0e-8 and 5.0e-8 * np.log(2.0) * 1e20 > 3.0e12
