#!/usr/bin/env python
"""bench.py -- electron-steps/s of the trapped-charge kinetics hot path on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c5|c3|c4|c1] [--replicas R] [--scaling weak|strong]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # CPU arm: the reference's algorithm on all host threads

One "step" = one pass of the hot path over one batch of replicas (seed the boxes, run every replica's schedule to the
end, fuse the ensemble histograms, all-reduce them across ranks; for the Optimizer workload c4: one population through
`mcl_objective`, objective values all-gathered).  Default workload at every N: BASELINE.json configs[1] (C2) --
isothermal hold then optical readout, N_e = 10^4 electrons per replica.

`--scaling weak` (default): `--replicas` replicas (c4: candidates) PER GPU; global replica ids are offset by rank, so
every replica of the job is distinct.  `--scaling strong`: `--replicas` is the size of the WHOLE job (default: c5 50 000
replicas = 10^8 electrons, c4 4096 candidates), split into contiguous blocks of global ids -- the result is identical for
any N.

Prints ONE JSON line (rank 0).  `value` = electron-steps of all ranks / max-over-ranks device time.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "electron-steps/sec (device-timed)"
UNIT = "electron-steps/s"

# Algorithmic cost of one electron-step: 3 SFU ops + 45 FP32/INT32 lane-ops
SFU_PER_ESTEP = 3.0
LANEOPS_PER_ESTEP = 45.0

DEFAULT_REPLICAS = {"weak": {"c2": 10_000, "c5": 6_250, "c1": 8, "c3": 2_560, "c4": 4_096},
                    "strong": {"c2": 10_000, "c5": 50_000, "c1": 8, "c3": 2_560, "c4": 4_096}}
LAB_OVERRIDES = ["exp_type_fp=TLlab", "physics_fp=lab_TL"]


def build_workload(name: str, replicas: int, electrons: int = 0, two_channel: bool = False):
    """Replica tables of a BASELINE config with `replicas` replicas (c4: a dict with the candidate matrix)."""
    from mcluminescence_b200 import workloads
    two_channel = two_channel or bool(os.environ.get("MCL_BENCH_TWO_CHANNEL"))
    if name == "c2":
        kw = {"physics_overrides": ["physics_fp.E_loc_2=1.0", "physics_fp.Retrap=0.3"]} if two_channel else {}
        return workloads.c2(n_replicas=replicas, n_e=electrons, **kw) if electrons > 0 else workloads.c2(n_replicas=replicas, **kw)
    if name == "c5":
        return workloads.c5(n_replicas=replicas)
    if name == "c1":
        return workloads.c1()
    if name == "c3":
        wl = workloads.c3(replicas_per_dose=max(1, -(-replicas // 10)))
        if len(wl["replicas"]) != replicas:                  # an even sample over the ten dose groups
            pick = np.linspace(0, len(wl["replicas"]) - 1, replicas).round().astype(int)
            wl["replicas"], wl["hist_group"] = wl["replicas"][pick], wl["hist_group"][pick]
        return wl
    if name == "c4":
        from mcluminescence_b200.config import compose
        return dict(name=f"C4 Optimizer inner loop: {replicas} Sobol candidates in DEFAULT_BOUNDS x 11 lab rows (tl_clbr), N_e=100",
                    population=workloads.c4_candidates(replicas, seed=4), cfg=compose(overrides=LAB_OVERRIDES), exp="tl_clbr",
                    hist=None)
    raise SystemExit(f"unknown workload {name}")


def c4_tables(wl, n_candidates: int):
    """Replica / segment tables of the first `n_candidates` candidates of a c4 workload (for the CPU legs)."""
    from mcluminescence_b200.config import DATA_DIR, initialize_runs
    from mcluminescence_b200.optimizer import cfg_with_params
    from mcluminescence_b200.replicas import LAB_CSV, LabTable
    lt = LabTable(*LAB_CSV[wl["exp"]], os.path.dirname(DATA_DIR))
    reps, segs = [], None
    for c in range(n_candidates):
        run = initialize_runs(cfg_with_params(wl["cfg"].deepcopy(), wl["population"][:, c]))[0]
        r, segs = lt.tables(run)
        reps.append(r)
    return np.concatenate(reps), segs, int(wl["cfg"]["exp_type_fp"]["steps"])


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for ts, line in self.rows:
            if ts < t0 or ts > t1 + 0.2:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax = float(f[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_leg(workload_name: str, threads: int, n_units: int, seed: int):
    """Time the CPU oracle (the C restatement of the reference's algorithm) on a bounded sample of the SAME replica
    shape: `n_units` replicas (c4: candidates).  Returns (electron-steps/s, electron-steps, seconds)."""
    from oracle import mcl_oracle as mo
    wl = build_workload(workload_name, n_units)
    obs = None
    if workload_name == "c4":
        reps, segs, max_steps = c4_tables(wl, n_units)
    else:
        reps, segs, max_steps = wl["replicas"][:n_units], wl["segments"], wl["max_steps"]
    t0 = time.perf_counter()
    res = mo.run(reps, segs, max_steps, seed=seed, obs_time=obs, parallel=True, threads=threads, trace=False)
    dt = time.perf_counter() - t0
    if workload_name != "c4" and res.rc != 0:
        raise RuntimeError(f"oracle failed with status {res.rc}")
    es = int(res.esteps.sum())
    return es / dt, es, dt


def cpu_sample_units(workload_name: str, threads: int) -> int:
    # about 10-30 s of CPU work per leg: one big box per thread, or many small ones
    return {"c2": 1, "c5": 6, "c1": 1, "c3": 4, "c4": 24}[workload_name] * max(threads, 1)


def numpy_reference_leg(seconds_budget: float = 20.0):
    """The UNMODIFIED NumPy reference (oracle/_ref, copied verbatim from /root/reference/src/class by
    oracle/ref_harness/make_ref.py at build time; absent => None) on a down-scaled C2 hold leg, one core."""
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.isfile(os.path.join(ref_dir, "src_class", "simulate.py")):
        return None
    try:
        out = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_harness", "time_ref.py"), str(seconds_budget)],
                             capture_output=True, text=True, timeout=40 + 4 * seconds_budget)
        return json.loads(out.stdout.strip().splitlines()[-1])
    except Exception as exc:  # noqa: BLE001
        return {"unavailable": f"{type(exc).__name__}: {exc}"}


def reference_arm(args):
    """--impl reference: the reference's own algorithm on the host cores, all host threads, bounded sample.  The
    NumPy reference itself cannot run these shapes (a dense 10^4 x 17 279 float64 matrix per replica, no optical leg,
    no multi-leg schedule), so the arm times its C port (oracle/mcl_oracle.c, pinned bit-for-bit against the reference's
    golden vectors); the NumPy reference's own rate on a down-scaled leg rides along in `cpu_baseline`."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import mcl_oracle as mo
    threads = mo.max_threads()
    n_units = cpu_sample_units(args.workload, threads)
    es_tot, secs = 0, 0.0
    for i in range(args.warmup + args.steps):
        v, es, dt = cpu_leg(args.workload, threads, n_units, seed=1000 + i)
        if i >= args.warmup:
            es_tot += es; secs += dt
    value = es_tot / secs
    unit_name = "candidates" if args.workload == "c4" else "replicas"
    sample = (f"{n_units} {unit_name} of the {args.workload.upper()} shape per step on {threads} host threads, {args.steps} steps")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(args.steps, 1),
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": build_workload(args.workload, args.replicas)["name"], "sample": sample},
        "cpu_baseline": {"value": value, "per_core": value / max(threads, 1), "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": sample, "numpy_reference": numpy_reference_leg(10.0)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c5", "c1", "c3", "c4"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--replicas", type=int, default=0,
                    help="replicas (c4: candidates) per GPU per step (weak) or of the whole job (strong); 0 = workload default")
    ap.add_argument("--electrons", type=int, default=0, help="tuning only: electrons per replica of the C2 shape")
    ap.add_argument("--seed", type=int, default=2)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg and the side measurements")
    args = ap.parse_args()
    if args.replicas <= 0:
        args.replicas = DEFAULT_REPLICAS[args.scaling][args.workload]

    if args.impl == "reference":
        reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from mcluminescence_b200 import engine, ensemble

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if world != args.gpus and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE {world}", file=sys.stderr)

    strong = args.scaling == "strong"
    wl = build_workload(args.workload, args.replicas, args.electrons)
    population = args.workload == "c4"
    peaks = engine.device_peaks()
    units_per_rank = (ensemble.shard_bounds(args.replicas, world, rank)[1] - ensemble.shard_bounds(args.replicas, world, rank)[0]) if strong else args.replicas

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def make_step(workload):
        def one_step(i):
            """(finish, device counters [esteps, steps, errors, final n_e], kernel ms or None); every step is a fresh
            ensemble: new Philox seed, same shapes."""
            seed = args.seed + 7919 * i
            if population:
                mse, es, kms = ensemble.run_population(workload["population"], workload["cfg"], workload["exp"], seed=seed, rank=rank,
                                                       world=world, shard=strong)
                counters = torch.tensor([es, 0, int((~np.isfinite(mse)).sum()), 0], dtype=torch.int64, device="cuda")
                return (lambda: mse), counters, kms
            finish, T = ensemble.run_ensemble(workload, seed=seed, rank=rank, world=world, shard=strong)
            return finish, T["counters"], None
        return one_step

    one_step = make_step(wl)
    for i in range(args.warmup):
        one_step(i)
    barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    stream = torch.cuda.current_stream()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    counters, kernel_ms_native = [], []
    wall0 = time.perf_counter()
    e_first, e_last = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = engine.launch_count()
    e_first.record(stream)
    for i in range(args.steps):
        ev[i][0].record(stream)
        finish, cnt, kms = one_step(args.warmup + i)
        ev[i][1].record(stream)
        counters.append(cnt)
        kernel_ms_native.append(kms)
    e_last.record(stream)
    launches = engine.launch_count() - launches0          # this rank's kernels inside the timed region (every rank launches the same number)
    barrier()
    wall1 = time.perf_counter()
    clocks = sampler.stop(wall0, wall1) if rank == 0 else None

    total_ms = e_first.elapsed_time(e_last)
    kernel_ms = [k if k is not None else a.elapsed_time(b) for (a, b), k in zip(ev, kernel_ms_native)]
    # counters were all-reduced inside the step: esteps is already the whole-job count
    esteps_job = int(sum(int(c[0].item()) for c in counters))
    errors = int(sum(int(c[2].item()) for c in counters))
    t_ms = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    max_ms = float(t_ms.item())
    value = esteps_job / (max_ms * 1e-3)

    # ---- e2e: the public API with HOST tables in and HOST results out, wall-clocked
    barrier()
    t0 = time.perf_counter()
    e2e_es = 0
    for i in range(args.steps):
        finish, cnt, _ = one_step(1000 + i)
        res = finish()                      # D2H of histograms + counters (c4: objective values), synchronises
        e2e_es += int(cnt[0].item()) if population else res.esteps
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = e2e_es / float(e2e_t.item())
    hist = wl["hist"]
    if population:
        h2d = int(wl["population"][:, :units_per_rank].nbytes + units_per_rank * 11 * 120 + 11 * 48)   # candidates + the replica / row tables built from them
        d2h = int(units_per_rank * 11 * 16 + units_per_rank * 8)
    else:
        h2d = int(wl["replicas"][:units_per_rank].nbytes + wl["segments"].nbytes)
        d2h = int((3 * hist.n_groups * hist.n_bins * 8 if hist is not None else 0) + 4 * 8)

    if rank == 0:
        # ---- roofline of the dominant kernel (philox_kernel: one launch per step and rank)
        n_sm = peaks["n_sm"]
        per_launch_es = esteps_job / max(args.steps, 1) / world
        kern_s = statistics.mean(kernel_ms) * 1e-3
        achieved = per_launch_es / kern_s
        sm_hz = (clocks["sm_mhz"] or peaks["sm_clock_mhz"]) * 1e6
        peak_measured = min(peaks["mufu_gops"] * 1e9 / SFU_PER_ESTEP, peaks["ffma_gops"] * 1e9 / LANEOPS_PER_ESTEP)
        peak_nominal = n_sm * 1.965e9 / max(SFU_PER_ESTEP / 16.0, LANEOPS_PER_ESTEP / 128.0)
        steps_job = int(sum(int(c[1].item()) for c in counters))
        hbm_bytes = 16.0 * steps_job / world / max(args.steps, 1)      # 16 B per (replica, step) record
        try:
            mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            hbm_peak, hbm_src = float(mp["hbm_gbs"]), "measured"
        except Exception:  # noqa: BLE001
            hbm_peak, hbm_src = 6650.0, "fallback"
        traffic, traffic_src = None, None
        for tf in ("r02_traffic.json", "r01_traffic_c2.json"):
            try:        # DRAM bytes of this kernel from a committed `ncu --set full` capture, scaled per replica: a profile constant, not measured in this run
                tr = json.load(open(os.path.join(ROOT, "profiles", tf)))
                tr = tr.get(args.workload, tr) if "workload" not in tr else tr
                if tr.get("workload") == args.workload:
                    traffic = (tr["dram_bytes_read"] + tr["dram_bytes_write"]) / tr["replicas_in_capture"] * units_per_rank
                    traffic_src = f"profiles/{tf}: ncu --set full capture of {tr['replicas_in_capture']} replicas (commit {tr.get('commit', 'round 1')}), scaled per replica; not measured in this run"
                    break
            except Exception:  # noqa: BLE001
                pass
        # A tighter bound from what the sweep of THIS workload actually issues per electron-step (SASS of the identical-channel
        # sweep: one Philox4x32-10 call = 20 IMAD.WIDE per four electrons, 2 MUFU.LG2 each, 20.25 issue slots in all; two
        # distinct channels: one call per two electrons).  The wide multiply is the scarce pipe (measured ~30 lanes/clk/SM).
        two_ch = bool(os.environ.get("MCL_BENCH_TWO_CHANNEL")) or args.workload == "c4"
        with_cb = args.workload in ("c3", "c4")      # lab_TL physics: a finite E_cb keeps the conduction-band term on (two more SFU ops per clock)
        mix = {"imad_wide": 10.0 if two_ch else 5.0, "mufu": 4.0 if with_cb else 2.0, "issue_slots": (30.0 if two_ch else 20.25) + (6.0 if with_cb else 0.0)}
        peak_mix = min(peaks["imad_gops"] * 1e9 / mix["imad_wide"], peaks["mufu_gops"] * 1e9 / mix["mufu"],
                       peaks["ffma_gops"] * 1e9 / mix["issue_slots"])
        roofline = {
            "bound": "sfu_fp32_issue", "achieved": achieved, "peak": peak_measured, "unit": UNIT,
            "frac": achieved / peak_measured, "traffic": traffic, "traffic_source": traffic_src,
            "kernel": "smallbox_kernel" if population else "philox_kernel",
            "model": f"{SFU_PER_ESTEP:g} SFU + {LANEOPS_PER_ESTEP:g} FP32/INT32 lane-ops per electron-step (SURVEY 8d)",
            "peak_source": "mcl_device_peaks microbenchmarks in this run (MUFU, FFMA issue)",
            "peak_nominal": peak_nominal, "frac_nominal": achieved / peak_nominal,
            "instruction_mix_bound": {"per_electron_step": mix, "peak": peak_mix, "frac": achieved / peak_mix,
                                      "note": "min over the measured pipe rates of what the sweep issues per electron-step; the wide-multiply pipe binds"},
            "pipe_peaks_gops": {k: peaks[k] for k in ("mufu_gops", "ffma_gops", "imad_gops", "lop3_gops")},
            "sm_mhz_during": sm_hz / 1e6,
            "hbm": {"achieved_gbs": hbm_bytes / kern_s / 1e9, "peak_gbs": hbm_peak, "peak_source": hbm_src,
                    "frac": hbm_bytes / kern_s / 1e9 / hbm_peak, "algorithmic_bytes_per_launch": hbm_bytes},
        }
        cpu, side = None, None
        if not args.no_cpu and world == 1:          # the CPU baseline rides on the N = 1 line only
            from oracle import mcl_oracle as mo
            threads = mo.max_threads()
            n_units = cpu_sample_units(args.workload, threads)
            v, es, dt = cpu_leg(args.workload, threads, n_units, seed=4242)
            unit_name = "candidates" if population else "replicas"
            cpu = {"value": v, "per_core": v / max(threads, 1), "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"{n_units} {unit_name} of the same shape on {threads} host threads, {dt:.1f} s",
                   "numpy_reference": numpy_reference_leg(15.0)}
            if args.workload == "c2" and world == 1 and not args.electrons:
                # side measurement: the same workload with two DISTINCT tunnelling channels (what every Optimizer candidate has)
                wl2 = build_workload("c2", min(args.replicas, 2960), two_channel=True)
                step2 = make_step(wl2)
                step2(0); torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream); _, c2, _ = step2(1); b.record(stream); torch.cuda.synchronize()
                side = {"c2_two_channel": {"value": int(c2[0].item()) / (a.elapsed_time(b) * 1e-3), "unit": UNIT,
                                           "replicas": len(wl2["replicas"]), "physics": "E_loc_2=1.0, Retrap=0.3"}}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": max_ms / max(args.steps, 1), "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["name"],
                       ("units_of_whole_job" if strong else "replicas_per_gpu_per_step"): int(args.replicas),
                       "l2": "working set per step (replica slabs) exceeds the 126 MB L2; fresh seed every step",
                       "rng": "philox4x32-10", "errors": errors},
            "clocks": clocks, "gpu_launches": launches * world,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "roofline": roofline, "cpu_baseline": cpu,
        }
        if side:
            line["side"] = side
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
