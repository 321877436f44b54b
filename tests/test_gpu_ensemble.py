"""GPU: multi-leg schedules, the fused integer histograms, sharding invariance and the batched objective."""
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


def two_leg_workload(R=96, n_e=200, n_bins=40):
    from mcluminescence_b200 import workloads
    from mcluminescence_b200.engine import AXIS_TIME_LOG, HistSpec
    wl = workloads.c2(n_replicas=R, n_e=n_e, n_bins=n_bins)
    wl["segments"]["duration"] = [200.0, 500.0]
    wl["segments"]["A_opt"] = [0.0, 50.0]
    wl["hist"] = HistSpec(axis=AXIS_TIME_LOG, n_bins=n_bins, lo=1e-2, hi=1e3, n_groups=2)
    wl["max_steps"] = 4000
    return wl


def rebin(wl, event, n_e, t, used, n_e0):
    """The kernel's histogram definition restated in NumPy from its own per-step trace."""
    h = wl["hist"]
    ratio = 10.0 ** ((np.log10(h.hi) - np.log10(h.lo)) / h.n_bins)
    edges = [h.lo]
    for _ in range(h.n_bins):
        edges.append(edges[-1] * ratio)                # same recurrence as the kernel
    edges = np.array(edges)
    n_seg = len(wl["segments"])
    ev = np.zeros((n_seg, h.n_bins), np.int64); occ = np.zeros_like(ev); occ2 = np.zeros_like(ev)
    durs = wl["segments"]["duration"]
    for r in range(event.shape[0]):
        seg, t_off, cursor, n_prev = 0, 0.0, 0, n_e0
        for i in range(used[r]):
            t_loc = t[r, i] - t_off
            while cursor <= h.n_bins and edges[cursor] <= t_loc:
                if cursor < h.n_bins:
                    occ[seg, cursor] += n_prev; occ2[seg, cursor] += n_prev * n_prev
                cursor += 1
            if event[r, i]:
                b = cursor - 1
                if 0 <= b < h.n_bins:
                    ev[seg, b] += 1
            n_prev = n_e[r, i]
            if t_loc >= durs[seg] and seg + 1 < n_seg:      # the leg ends with the step that passes its duration
                seg += 1; t_off = t[r, i]; cursor = 0
    return ev, occ, occ2


def test_fused_histograms_equal_rebinned_traces(gpu):
    from mcluminescence_b200 import engine
    wl = two_leg_workload()
    out = engine.run_replicas(wl["replicas"], wl["segments"], wl["max_steps"], seed=11, hist=wl["hist"],
                              trace=True, sync=True)
    out.raise_on_error()
    ev, occ, occ2 = rebin(wl, out.event, out.n_e, out.t, out.steps_used, int(wl["replicas"]["n_e0"][0]))
    assert np.array_equal(out.hist_events, ev)
    assert np.array_equal(out.hist_occ, occ)
    assert np.array_equal(out.hist_occ_sq, occ2)
    assert ev[0].sum() > 0 and ev[1].sum() > 0          # both legs produced luminescence
    assert int(out.hist_events.sum()) <= int(out.event.sum())


def test_two_leg_schedule_matches_oracle_statistically(gpu):
    from mcluminescence_b200 import engine
    from oracle import mcl_oracle as mo
    wl = two_leg_workload(R=256)
    out = engine.run_replicas(wl["replicas"], wl["segments"], wl["max_steps"], seed=5, sync=True)
    out.raise_on_error()
    ref = mo.run(wl["replicas"], wl["segments"], wl["max_steps"], seed=31, parallel=True)
    assert ref.rc == 0
    durs = wl["segments"]["duration"]

    def per_leg(event, t, used):
        a = np.zeros((event.shape[0], 2))
        for r in range(event.shape[0]):
            tt, e = t[r, :used[r]], event[r, :used[r]]
            k = int(np.searchsorted(tt, durs[0], side="left")) + 1     # steps of leg 0 (incl. the overshooting one)
            a[r] = e[:k].sum(), e[k:].sum()
        return a

    g, o = per_leg(out.event, out.t, out.steps_used), per_leg(ref.event, ref.t, ref.steps_used)
    for leg in range(2):
        se = np.sqrt(g[:, leg].var(ddof=1) / len(g) + o[:, leg].var(ddof=1) / len(o))
        assert abs(g[:, leg].mean() - o[:, leg].mean()) <= 4 * se, (leg, g[:, leg].mean(), o[:, leg].mean())
    se = np.sqrt(out.final_n_e.var(ddof=1) / len(g) + ref.final_n_e.var(ddof=1) / len(o))
    assert abs(out.final_n_e.mean() - ref.final_n_e.mean()) <= 4 * se


def test_ensemble_is_invariant_to_sharding(gpu):
    """Integer histograms + (seed, global replica id) keyed streams: any split sums to the same result."""
    from mcluminescence_b200 import ensemble
    wl = two_leg_workload(R=60)
    full, _ = ensemble.run_ensemble(wl, seed=3)
    full = full()
    parts = []
    for rank in range(3):
        fin, _ = ensemble.run_ensemble(wl, seed=3, rank=rank, world=3, reduce=False)
        parts.append(fin())
    assert np.array_equal(sum(p.hist_events for p in parts), full.hist_events)
    assert np.array_equal(sum(p.hist_occ for p in parts), full.hist_occ)
    assert np.array_equal(sum(p.hist_occ_sq for p in parts), full.hist_occ_sq)
    assert sum(p.esteps for p in parts) == full.esteps and full.errors == 0


def test_objective_batched_equals_single_candidate_calls(gpu, capsys):
    from mcluminescence_b200 import optimizer
    from mcluminescence_b200.config import compose
    from mcluminescence_b200.workloads import c4_candidates
    cfg = compose(overrides=helpers.LAB_OVERRIDES)
    P = c4_candidates(8, seed=4)                       # [10, 8]
    for exp in ("tl_clbr", "iso"):
        batch = optimizer.objective_batched(P, cfg, exp, seed=21)
        assert batch.shape == (8,) and np.all(np.isfinite(batch)) and np.all(batch >= 0)
        for c in (0, 3, 7):
            one = optimizer.objective(P[:, c], cfg, exp, seed=21, candidate_id=c)
            assert one == batch[c], (exp, c)
    with pytest.raises(ValueError):
        optimizer.objective_batched(P, cfg, "nope")
    capsys.readouterr()


def test_objective_statistics_match_oracle(gpu, capsys):
    """Mean objective over many independent evaluations, GPU Philox vs CPU oracle, at the YAML defaults."""
    from mcluminescence_b200 import optimizer
    from mcluminescence_b200.config import compose, initialize_runs
    from mcluminescence_b200.replicas import LAB_CSV, LabTable
    from oracle import mcl_oracle as mo
    cfg = compose(overrides=helpers.LAB_OVERRIDES)
    run = initialize_runs(cfg)[0]
    p0 = np.array([run.exp_type_fp.rho_prime, run.physics_fp.E_cb, run.physics_fp.E_loc_1, run.physics_fp.E_loc_2,
                   run.physics_fp.D0, run.physics_fp.s, run.physics_fp.b, run.physics_fp.alpha,
                   run.exp_type_fp.holes, run.physics_fp.Retrap], dtype=float)
    M = 64
    g = optimizer.objective_batched(np.tile(p0[:, None], (1, M)), cfg, "tl_clbr", seed=8)
    lt = LabTable(*LAB_CSV["tl_clbr"], helpers.DATA_ROOT)
    reps1, segs = lt.tables(run)
    ref = mo.run(np.tile(reps1, M), segs, int(run.exp_type_fp.steps), seed=600, parallel=True, trace=False)
    o = np.array([lt.mse(run.exp_type_fp.N_e, ref.final_n_e[m * 11:(m + 1) * 11])[1] for m in range(M)])
    se = np.sqrt(g.var(ddof=1) / M + o.var(ddof=1) / M)
    assert abs(g.mean() - o.mean()) <= 4 * se, (g.mean(), o.mean(), se)
    capsys.readouterr()


@pytest.mark.parametrize("which", ["c2", "c5"])
def test_full_size_replicas_conserve_charge(gpu, which):
    """BASELINE-sized replicas (10^4 / 2000 electrons): size-independent invariants of the kernel's own
    outputs -- every recombination removes exactly one electron, histograms account for every event inside
    the axis, occupancy is non-increasing without a dose, and the count is independent of the launch shape."""
    from mcluminescence_b200 import engine, workloads
    wl = workloads.c2(n_replicas=24) if which == "c2" else workloads.c5(n_replicas=96)
    out = engine.run_replicas(wl["replicas"], wl["segments"], wl["max_steps"], seed=17, hist=wl["hist"],
                              trace=True, sync=True)
    out.raise_on_error()
    n0 = wl["replicas"]["n_e0"].astype(np.int64)
    events = np.array([out.event[r, :out.steps_used[r]].sum() for r in range(len(n0))])
    assert np.array_equal(events + out.final_n_e, n0)                   # charge conservation, no fills
    for r in range(0, len(n0), 7):
        n = out.steps_used[r]
        ne = out.n_e[r, :n]
        assert np.all(np.diff(ne) <= 0) and np.all(np.diff(ne) >= -1)
        assert np.all(np.diff(out.t[r, :n]) >= 0)
        assert np.array_equal(np.concatenate([[n0[r]], ne[:-1]]) - ne, out.event[r, :n])
    assert int(out.hist_events.sum()) <= int(events.sum())
    assert int(out.esteps.sum()) == int(sum((out.n_e[r, :out.steps_used[r]] + out.event[r, :out.steps_used[r]]).sum()
                                            for r in range(len(n0))))
    # occupancy histogram is a sum of non-increasing step functions
    occ = out.hist_occ
    for row in range(occ.shape[0]):
        nz = occ[row][occ[row] > 0]
        assert np.all(np.diff(nz) <= 0)
    # same replicas, forced onto a different CTA width: identical traces
    import os
    from mcluminescence_b200 import _native
    os.environ["MCL_PHILOX_NT"] = "128" if which == "c2" else "32"
    try:
        alt = engine.run_replicas(wl["replicas"][:6], wl["segments"], wl["max_steps"], seed=17, trace=True, sync=True)
    finally:
        del os.environ["MCL_PHILOX_NT"]
    assert np.array_equal(alt.event, out.event[:6]) and np.array_equal(alt.n_e, out.n_e[:6])
    assert np.array_equal(alt.t, out.t[:6])


def test_objective_statistics_across_the_parameter_bounds(gpu, capsys):
    """Sobol candidates spanning DEFAULT_BOUNDS (alpha up to 5e10, rho' down to 1e-8, E_cb ~2 eV: the
    regimes where alpha*r is in the hundreds and the conduction-band channel dominates): the mean final
    fill of every lab row, GPU Philox vs CPU oracle, over 48 independent evaluations per candidate."""
    from mcluminescence_b200 import optimizer
    from mcluminescence_b200.config import compose, initialize_runs
    from mcluminescence_b200.optimizer import cfg_with_params
    from mcluminescence_b200.replicas import LAB_CSV, LabTable
    from mcluminescence_b200.workloads import c4_candidates
    from oracle import mcl_oracle as mo
    from mcluminescence_b200 import engine
    cfg = compose(overrides=helpers.LAB_OVERRIDES)
    P = c4_candidates(6, seed=11)
    lt = LabTable(*LAB_CSV["tl_clbr"], helpers.DATA_ROOT)
    M = 48
    for c in range(P.shape[1]):
        run = initialize_runs(cfg_with_params(cfg.deepcopy(), P[:, c]))[0]
        reps1, segs = lt.tables(run)
        reps = np.tile(reps1, M)
        steps = int(run.exp_type_fp.steps)
        out = engine.run_replicas(reps, segs, steps, seed=50 + c, trace=False, sync=True)
        ref = mo.run(reps, segs, steps, seed=900 + c, parallel=True, trace=False)
        if ref.rc != 0 or np.any(out.status != 0):
            # a candidate the reference itself cannot finish within `steps`: both must agree on that
            assert ref.rc != 0 and np.any(out.status != 0), (c, ref.rc, out.status[out.status != 0][:3])
            continue
        g = out.final_n_e.reshape(M, -1).astype(float)
        o = ref.final_n_e.reshape(M, -1).astype(float)
        for k in range(g.shape[1]):
            se = np.sqrt(g[:, k].var(ddof=1) / M + o[:, k].var(ddof=1) / M)
            assert abs(g[:, k].mean() - o[:, k].mean()) <= 4.5 * se + 0.35, (c, k, g[:, k].mean(), o[:, k].mean(), se)
    capsys.readouterr()


def test_32_bit_nearest_slot_path_gives_identical_results(gpu):
    """Boxes with more than 65534 hole slots keep 32-bit nearest-hole slots in shared memory; forcing that
    instantiation on small boxes must not change a single record (simulate legs, histograms, lab rows with fills)."""
    import os
    from mcluminescence_b200 import engine
    from mcluminescence_b200.config import compose, initialize_runs
    from mcluminescence_b200.replicas import LAB_CSV, LabTable
    wl = two_leg_workload(R=40, n_e=600)
    run = initialize_runs(compose(overrides=helpers.LAB_OVERRIDES))[0]
    lt = LabTable(*LAB_CSV["iso"], helpers.DATA_ROOT)
    lab_reps, lab_segs = lt.tables(run)

    os.environ["MCL_SMALLBOX"] = "0"                  # lab rows on the block kernel, which is what has two slot widths

    def both():
        a = engine.run_replicas(wl["replicas"], wl["segments"], wl["max_steps"], seed=9, hist=wl["hist"], trace=True, sync=True)
        b = engine.run_replicas(lab_reps, lab_segs, 20000, seed=9, obs_time=lt.obs_time, trace=True, sync=True)
        return a, b
    a16, b16 = both()
    os.environ["MCL_PHILOX_NEAR32"] = "1"
    try:
        a32, b32 = both()
    finally:
        del os.environ["MCL_PHILOX_NEAR32"]
        del os.environ["MCL_SMALLBOX"]
    for x, y in ((a16, a32), (b16, b32)):
        x.raise_on_error(); y.raise_on_error()
        assert np.array_equal(x.event, y.event) and np.array_equal(x.n_e, y.n_e) and np.array_equal(x.t, y.t)
        assert np.array_equal(x.esteps, y.esteps)
    assert np.array_equal(a16.hist_events, a32.hist_events) and np.array_equal(a16.hist_occ, a32.hist_occ)
    assert np.array_equal(b16.obs_n_e, b32.obs_n_e)


@pytest.mark.parametrize("which", ["c2", "c5"])
def test_scan_skip_bitmaps_never_hide_a_cached_electron(gpu, which):
    """Self-check mode: events whose post-event scan the sharing bitmaps would skip are scanned anyway and any hit
    is reported as MCL_ERR_INTERNAL.  Also: skipping changes no record."""
    import os
    from mcluminescence_b200 import engine, workloads
    wl = workloads.c2(n_replicas=32, n_e=6000) if which == "c2" else workloads.c5(n_replicas=64)
    runs = {}
    for mode in ("0", "1", "2"):
        os.environ["MCL_PHILOX_SHARE_BM"] = mode
        try:
            runs[mode] = engine.run_replicas(wl["replicas"], wl["segments"], wl["max_steps"], seed=23, trace=True, sync=True)
        finally:
            del os.environ["MCL_PHILOX_SHARE_BM"]
        assert not runs[mode].status.any(), (mode, runs[mode].status[runs[mode].status != 0][:4])
    for mode in ("1", "2"):
        assert np.array_equal(runs[mode].event, runs["0"].event) and np.array_equal(runs[mode].n_e, runs["0"].n_e)
        assert np.array_equal(runs[mode].t, runs["0"].t)


def test_specialised_step_loop_changes_no_result(gpu):
    """Legs of the simulate protocol without a dose and without per-step records run a specialised copy of the step
    loop (protocol / dose / trace / fill-mode branches compiled out; the general loop takes over when the filling
    clock could matter).  Histograms, final occupancies, step and electron-step counts must equal those of the
    general loop -- selected with `MCL_PHILOX_FAST=0`, and, independently, by asking for the per-step trace."""
    import os
    from mcluminescence_b200 import engine, workloads
    from tests.test_gpu_philox import CASES, ensemble_tables
    jobs = [workloads.c2(n_replicas=12), workloads.c5(n_replicas=300), two_leg_workload(R=40, n_e=600),
            workloads.c2(n_replicas=6, n_e=5000, physics_overrides=["physics_fp.E_loc_2=1.0", "physics_fp.Retrap=0.3"])]
    # ramps (conduction-band channel on and off) and, last, boxes that run almost empty: their final steps have no clock
    # below 5e12 s, which is where the specialised loop hands the step over to the general one
    for name in ("tl_ramp", "cb_channel"):
        reps, segs, steps = ensemble_tables(CASES[name][0], 64)
        jobs.append(dict(name=name, replicas=reps, segments=segs, max_steps=steps, hist=None))
    empty = two_leg_workload(R=16, n_e=40)
    empty["segments"]["duration"] = [1e6, 1e9]
    empty["max_steps"] = 4000
    jobs.append(empty)
    for wl in jobs:
        def go(trace):
            return engine.run_replicas(wl["replicas"], wl["segments"], wl["max_steps"], seed=43, hist=wl.get("hist"),
                                       trace=trace, sync=True)
        fast = go(False)
        with_trace = go(True)
        os.environ["MCL_PHILOX_FAST"] = "0"
        try:
            general = go(False)
        finally:
            del os.environ["MCL_PHILOX_FAST"]
        for other in (with_trace, general):
            assert np.array_equal(fast.status, other.status), wl["name"]
            assert np.array_equal(fast.steps_used, other.steps_used), wl["name"]
            assert np.array_equal(fast.final_n_e, other.final_n_e) and np.array_equal(fast.esteps, other.esteps), wl["name"]
            if wl.get("hist") is not None:
                assert np.array_equal(fast.hist_events, other.hist_events), wl["name"]
                assert np.array_equal(fast.hist_occ, other.hist_occ), wl["name"]
    # the last job ends with a handful of electrons whose next clock is beyond 5e12 s: the hand-over step
    assert 0 < int(fast.final_n_e.max()) < int(jobs[-1]["replicas"]["n_e0"][0]) // 2


def test_pipelined_step_loop_changes_no_result(gpu):
    """Isothermal legs of wide CTAs run the specialised loop PIPELINED: a sweep team computes the clocks of step k+1 while
    one warp decides step k, electrons cached on a dead hole are re-targeted lazily (only when they win a step with their
    stale lower bound), and steps the pipeline cannot take (compaction due, end of a leg, filling clock, a box running
    empty) are handed back and run in order.  `MCL_PHILOX_PIPE=0` runs the same legs in order: histograms, final
    occupancies, step and electron-step counts and status codes must be equal, replica for replica."""
    import os
    from mcluminescence_b200 import engine, workloads
    two = ["physics_fp.E_loc_2=1.0", "physics_fp.Retrap=0.3"]
    capped = workloads.c2(n_replicas=6, n_e=3000, n_bins=50)
    capped["segments"]["dt_cap"] = [0.25, 1e20]                 # most steps of the hold leg end at the cap: no event
    capped["max_steps"] = 30000
    ramp_cb = workloads.c5(n_replicas=32)
    # the kernel's (very conservative) test for "the conduction-band term cannot matter" compares it with the tunnelling rate at
    # the box diagonal, alpha r = 910 here: E_cb = 35 eV passes it below ~600 K and fails it above, i.e. in the middle of this ramp
    # (273 K -> 1073 K) the pipeline hands the leg back for good.  (basicTL12 ships E_cb = 1000 eV: never on.)
    ramp_cb["replicas"]["E_cb"] = 35.0
    short = workloads.c2(n_replicas=6, n_e=3000, n_bins=50)
    short["max_steps"] = 700                                    # MCL_ERR_STEPS in the middle of the first leg
    jobs = [("c2, 512 threads", workloads.c2(n_replicas=12), None),
            ("c2, 256 threads", workloads.c2(n_replicas=6), 256),
            ("c2 two channels", workloads.c2(n_replicas=6, n_e=5000, physics_overrides=two), 256),
            ("c2 2000 electrons, 128 threads", workloads.c2(n_replicas=24, n_e=2000, n_bins=50), 128),
            # conduction-band channel on (every clock takes the 4-SFU form); the boxes run nearly empty, so the legs end in order
            ("c2 conduction band", workloads.c2(n_replicas=8, n_e=3000, n_bins=50, physics_overrides=["physics_fp.E_cb=1.5"]), 256),
            ("c2 conduction band + two channels", workloads.c2(n_replicas=8, n_e=3000, n_bins=50, physics_overrides=["physics_fp.E_cb=1.5"] + two), 128),
            ("c2 step cap", capped, 256), ("c2 out of steps", short, 256),
            # ramps: identical channels and no conduction-band term -> the temperature is one uniform offset and the sweep team
            # does not need it; with a finite E_cb the conduction-band term switches on during the ramp and the leg ends in order
            ("c5 ramp, 128 threads", workloads.c5(n_replicas=48), 128), ("c5 ramp, 256 threads", workloads.c5(n_replicas=24), 256),
            ("ramp into the conduction band", ramp_cb, 128)]
    took_pipeline = False
    for name, wl, nt in jobs:
        def go(pipe):
            os.environ["MCL_PHILOX_PIPE"] = pipe
            if nt:
                os.environ["MCL_PHILOX_NT"] = str(nt)
            try:
                return engine.run_replicas(wl["replicas"], wl["segments"], wl["max_steps"], seed=43, hist=wl["hist"], trace=False, sync=True)
            finally:
                os.environ.pop("MCL_PHILOX_PIPE", None); os.environ.pop("MCL_PHILOX_NT", None)
        a, b = go("1"), go("0")
        for k in ("status", "steps_used", "final_n_e", "esteps", "hist_events", "hist_occ"):
            assert np.array_equal(getattr(a, k), getattr(b, k)), f"{name}: {k}"
        if name == "c2 out of steps":
            assert np.all(a.status != 0)
        else:
            a.raise_on_error()
        if name == "c2 conduction band":
            assert int(a.final_n_e.max()) < 1024               # below 4 x 256: the legs cannot have ended inside the pipeline
        took_pipeline |= bool(a.steps_used.min() > 1000)
    assert took_pipeline


def test_shared_memory_slab_changes_no_result(gpu, capsys):
    """Small boxes (the Optimizer path) keep their hole table, cell tables and electron coordinates in shared memory;
    `MCL_PHILOX_SMEM_SLAB=0` keeps them in the HBM slab.  Same algorithm: every objective value and electron-step
    count must be identical -- TL rows with fills and regrids, isothermal experiments with observation times."""
    import os
    from mcluminescence_b200 import optimizer
    from mcluminescence_b200.config import compose
    from mcluminescence_b200.workloads import c4_candidates
    cfg = compose(overrides=helpers.LAB_OVERRIDES)
    P = c4_candidates(64, seed=9)
    os.environ["MCL_SMALLBOX"] = "0"                  # this test is about the block kernel's two slab placements
    try:
        _slab_equality(P, cfg, optimizer)
    finally:
        del os.environ["MCL_SMALLBOX"]
    capsys.readouterr()


def _slab_equality(P, cfg, optimizer):
    import os
    for exp in ("tl_clbr", "iso"):
        for extra in (None, "-92"):
            if extra is not None:
                os.environ["MCL_PHILOX_FILL_EXTRA"] = extra
            try:
                a, ea = optimizer.objective_batched(P, cfg, exp, seed=31, return_esteps=True)
                os.environ["MCL_PHILOX_SMEM_SLAB"] = "0"
                try:
                    b, eb = optimizer.objective_batched(P, cfg, exp, seed=31, return_esteps=True)
                finally:
                    del os.environ["MCL_PHILOX_SMEM_SLAB"]
            finally:
                os.environ.pop("MCL_PHILOX_FILL_EXTRA", None)
            assert np.array_equal(a, b) and ea == eb, (exp, extra)


def test_smallbox_kernel_traces_are_consistent(gpu):
    """The one-warp-per-replica kernel of the Optimizer path: its per-step records must be consistent with its summary
    outputs (charge bookkeeping with fills and recombinations, electron-steps, observation crossings), must not depend
    on whether records are requested, nor on the batch a replica runs in."""
    from mcluminescence_b200 import engine
    from mcluminescence_b200.config import compose, initialize_runs
    from mcluminescence_b200.replicas import LAB_CSV, LabTable
    run = initialize_runs(compose(overrides=helpers.LAB_OVERRIDES))[0]
    for exp in ("tl_clbr", "iso"):
        lt = LabTable(*LAB_CSV[exp], helpers.DATA_ROOT)
        reps1, segs = lt.tables(run)
        M, n_rows = 8, len(reps1)
        reps = np.tile(reps1, M)
        obs_time = np.tile(lt.obs_time, M)
        if lt.obs_time.size:
            for m in range(M):
                reps["obs_begin"][m * n_rows:(m + 1) * n_rows] += m * len(lt.obs_time)
        tr = engine.run_replicas(reps, segs, 20000, seed=61, obs_time=obs_time, trace=True, sync=True)
        fin = engine.run_replicas(reps, segs, 20000, seed=61, obs_time=obs_time, trace=False, sync=True)
        tr.raise_on_error(); fin.raise_on_error()
        assert np.array_equal(tr.final_n_e, fin.final_n_e) and np.array_equal(tr.esteps, fin.esteps)
        assert np.array_equal(tr.steps_used, fin.steps_used) and np.array_equal(tr.obs_n_e, fin.obs_n_e)
        if not lt.obs_time.size:                        # the same replicas in another batch: keyed by global replica id
            part = engine.run_replicas(reps[n_rows:3 * n_rows], segs, 20000, seed=61, replica_id0=n_rows, trace=False, sync=True)
            assert np.array_equal(part.final_n_e, fin.final_n_e[n_rows:3 * n_rows])
            assert np.array_equal(part.esteps, fin.esteps[n_rows:3 * n_rows])
        for r in range(len(reps)):
            n = int(tr.steps_used[r])
            ne, ev, t = tr.n_e[r, :n].astype(np.int64), tr.event[r, :n], tr.t[r, :n]
            before = np.concatenate([[int(reps["n_e0"][r])], ne[:-1]])
            assert np.all((ne - before == 1) | ((ne - before == -1) & (ev == 1)))      # a step adds a pair or removes one
            assert np.all(ev[ne - before == 1] == 0) and np.all(np.diff(t) >= 0)
            assert int(before.sum()) == int(tr.esteps[r]) and int(ne[-1]) == int(tr.final_n_e[r])
            assert ne.max() <= int(reps["N_e"][r]) + 1
