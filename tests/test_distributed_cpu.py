"""World-size-2 checks of the multi-GPU host logic on CPU (gloo): sharding and the single all-reduce."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mcluminescence_b200.ensemble import allreduce_ensemble, shard_bounds
from mcluminescence_b200.engine import HistSpec


def test_shard_bounds_partition_the_ensemble():
    for n in (0, 1, 7, 8, 10_000, 50_001):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        hist = HistSpec(axis=1, n_bins=16, lo=1e-2, hi=1e2, n_groups=2)
        g = torch.Generator().manual_seed(100 + rank)
        T = {k: torch.randint(0, 1000, (2, 16), generator=g, dtype=torch.int64)
             for k in ("hist_events", "hist_occ", "hist_occ_sq")}
        counters = torch.tensor([10 + rank, 20 + rank, rank, 5], dtype=torch.int64)
        mine = {k: v.clone() for k, v in T.items()}
        red = allreduce_ensemble(T, hist, counters)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), counters=red.numpy(),
                 **{k: v.numpy() for k, v in T.items()}, **{"mine_" + k: v.numpy() for k, v in mine.items()})
        # counters-only variant (no histogram requested)
        c2 = allreduce_ensemble({}, None, torch.tensor([1, 2, 3, 4 + rank], dtype=torch.int64))
        assert c2.tolist() == [world, 2 * world, 3 * world, sum(4 + r for r in range(world))]
    finally:
        dist.destroy_process_group()


def test_single_allreduce_sums_histograms_and_counters(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r = [np.load(tmp_path / f"rank{k}.npz") for k in range(world)]
    for k in ("hist_events", "hist_occ", "hist_occ_sq"):
        total = r[0]["mine_" + k] + r[1]["mine_" + k]
        assert np.array_equal(r[0][k], total) and np.array_equal(r[1][k], total)
    assert r[0]["counters"].tolist() == [21, 41, 1, 10] == r[1]["counters"].tolist()
