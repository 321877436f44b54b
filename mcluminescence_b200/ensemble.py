"""Replica ensembles across GPUs: shard, run, one NCCL all-reduce of the integer histograms.

Replicas never interact (``src/class/simulate.py:36,46``: independent loops), so ranks take
contiguous blocks of global replica ids and run them with no data-path collective.  The only
exchange is the final ensemble sum: every rank's fused integer histograms (events, occupancy,
occupancy squared -- a few KB) are all-reduced once, plus a handful of scalar counters.
Because the Philox streams are keyed by (seed, global replica id) and the histograms are
integers, the result is identical for any number of ranks.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, Optional

import numpy as np

from . import engine
from .engine import HistSpec
from .replicas import MODE_PHILOX


def shard_bounds(n_total: int, world: int, rank: int):
    """Contiguous block [lo, hi) of global replica ids owned by ``rank`` (sizes differ by <= 1)."""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_ensemble(T: Dict[str, Any], hist: Optional[HistSpec], counters, group=None):
    """The ONE collective of the path: sum the integer histograms and the scalar counters over ranks.

    Everything is packed into a single int64 vector so that exactly one all-reduce (NCCL on GPUs,
    gloo in the CPU tests) is issued per ensemble.  Updates ``T`` in place, returns the counters.
    """
    import torch
    import torch.distributed as dist
    if hist is not None:
        packed = torch.cat([T["hist_events"].reshape(-1), T["hist_occ"].reshape(-1),
                            T["hist_occ_sq"].reshape(-1), counters])
    else:
        packed = counters.clone()
    dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    if hist is not None:
        n = hist.n_groups * hist.n_bins
        T["hist_events"] = packed[0:n].reshape(hist.n_groups, hist.n_bins)
        T["hist_occ"] = packed[n:2 * n].reshape(hist.n_groups, hist.n_bins)
        T["hist_occ_sq"] = packed[2 * n:3 * n].reshape(hist.n_groups, hist.n_bins)
        return packed[3 * n:]
    return packed


@dataclass
class EnsembleResult:
    hist_events: np.ndarray      # [rows, bins] int64, summed over all replicas of all ranks
    hist_occ: np.ndarray         # [rows, bins] int64  sum of n_e at each bin's left edge
    hist_occ_sq: np.ndarray      # [rows, bins] int64  sum of n_e^2
    n_replicas: int              # global
    esteps: int                  # global electron-steps
    steps: int                   # global steps
    errors: int                  # replicas that ended with a non-zero status
    final_n_e_sum: int

    def mean_occupancy(self, replicas_per_row):
        return self.hist_occ / np.asarray(replicas_per_row, dtype=np.float64)[:, None]

    def std_occupancy(self, replicas_per_row):
        n = np.asarray(replicas_per_row, dtype=np.float64)[:, None]
        m = self.hist_occ / n
        return np.sqrt(np.maximum(self.hist_occ_sq / n - m * m, 0.0))


def run_ensemble(workload: Dict[str, Any], *, seed: int, rank: int = 0, world: int = 1,
                 group=None, replica_id0: int = 0, shard: bool = True, device=None,
                 reduce: bool = True):
    """Run a workload dict (see ``workloads``) on this rank's shard and reduce across ranks.

    ``shard=True``: the workload's replicas are the GLOBAL ensemble and this rank runs its block.
    ``shard=False``: the replicas given are this rank's own (weak scaling); global ids are
    ``replica_id0 + rank * len(replicas) + r``.
    Returns ``(EnsembleResult, device_tensors)``; the tensors stay on the GPU for callers that time
    the kernel separately from the read-back.
    """
    torch = engine._torch()
    reps, segs = workload["replicas"], workload["segments"]
    hist: Optional[HistSpec] = workload.get("hist")
    grp = workload.get("hist_group")
    if shard:
        lo, hi = shard_bounds(len(reps), world, rank)
        my = reps[lo:hi]
        my_grp = grp[lo:hi] if grp is not None else None
        id0 = replica_id0 + lo
        n_global = len(reps)
    else:
        my, my_grp = reps, grp
        id0 = replica_id0 + rank * len(reps)
        n_global = len(reps) * world
    out = engine.run_replicas(my, segs, workload["max_steps"], mode=MODE_PHILOX, seed=seed,
                              replica_id0=id0, trace=False, hist=hist, hist_group=my_grp,
                              device=device)
    T = out.tensors
    counters = torch.stack([T["esteps"].sum(), T["steps_used"].sum().to(torch.int64),
                            (T["status"] != 0).sum().to(torch.int64), T["final_n_e"].sum().to(torch.int64)])
    if reduce and world > 1:
        counters = allreduce_ensemble(T, hist, counters, group)
    T["counters"] = counters

    def finish() -> EnsembleResult:
        c = counters.cpu().numpy()
        z = np.zeros((1, 1), np.int64)
        return EnsembleResult(
            hist_events=T["hist_events"].cpu().numpy() if hist is not None else z,
            hist_occ=T["hist_occ"].cpu().numpy() if hist is not None else z,
            hist_occ_sq=T["hist_occ_sq"].cpu().numpy() if hist is not None else z,
            n_replicas=n_global, esteps=int(c[0]), steps=int(c[1]), errors=int(c[2]), final_n_e_sum=int(c[3]))

    return finish, T


def run_population(P: np.ndarray, cfg, exp: str, *, seed: int, rank: int = 0, world: int = 1, group=None,
                   shard: bool = True, candidate_id0: int = 0):
    """Optimizer population across GPUs (``src/class/optimizer.py:104-111`` maps ``objective`` over a process pool):
    every rank evaluates its block of candidates with one ``mcl_objective`` call, then ONE all-gather returns the
    objective values to every rank (plus an all-reduce of two counters).  Candidate ``c`` uses the Philox streams of
    global id ``candidate_id0 + c`` whatever the number of ranks.

    ``shard=True``: ``P[10, S]`` is the global population, this rank takes its block.  ``shard=False``: ``P`` is this
    rank's own population (weak scaling), global ids offset by rank.
    Returns ``(mse_global[S_total], electron_steps_global, kernel_ms_this_rank)``.
    """
    from . import _native, optimizer
    torch = engine._torch()
    S = int(P.shape[1])
    if shard:
        lo, hi = shard_bounds(S, world, rank)
        mine, id0, S_total = P[:, lo:hi], candidate_id0 + lo, S
    else:
        mine, id0, S_total = P, candidate_id0 + rank * S, S * world
    if mine.shape[1] > 0:
        mse, es = optimizer.objective_batched(np.ascontiguousarray(mine), cfg, exp, seed=seed, candidate_id0=id0,
                                              return_esteps=True)
        kernel_ms = float(_native.load().mcl_objective_last_kernel_ms())
    else:
        mse, es, kernel_ms = np.zeros(0), 0, 0.0
    if world == 1:
        return mse, es, kernel_ms
    import torch.distributed as dist
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    counts = [shard_bounds(S, world, r)[1] - shard_bounds(S, world, r)[0] for r in range(world)] if shard else [S] * world
    pad = max(counts)
    buf = torch.zeros(pad, dtype=torch.float64, device=dev)
    buf[:mse.size] = torch.as_tensor(mse, dtype=torch.float64)
    gathered = [torch.zeros(pad, dtype=torch.float64, device=dev) for _ in range(world)]
    dist.all_gather(gathered, buf, group=group)
    tot = torch.tensor([es], dtype=torch.int64, device=dev)
    dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=group)
    out = np.concatenate([g[:n].cpu().numpy() for g, n in zip(gathered, counts)])
    return out, int(tot.item()), kernel_ms
