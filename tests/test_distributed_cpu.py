"""world_size-2/3 gloo tests (CPU) of the N > 1 host path: deterministic cost sharding, padded all-gather of the
packed vectors, scatter into the dense matrices.  The per-pair values are a known function of (i, j) here; on the
GPU box tests/test_gpu_pairs.py::test_shards_partition_the_pairs checks the real kernels shard by shard."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from caretta_b200 import distributed as D
from caretta_b200 import engine


def _fake(pi, pj, f):
    return (pi.astype(np.float64) * 131 + pj * 7 + f * 0.5).astype(np.float32)


def _worker(rank, world, port, offsets, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = len(offsets) - 1
        shards, pad = D.shard_layout(offsets, world)
        pi, pj = shards[rank]
        local = torch.zeros(3 * pad, dtype=torch.float32)
        for f in range(3):
            local[f * pad:f * pad + len(pi)] = torch.from_numpy(_fake(pi, pj, f))
        gathered = D.gather_packed(local, world)
        mats = D.scatter_to_matrices(gathered, shards, pad, n)
        q.put((rank, mats["score"], mats["rmsd"], mats["tm"]))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 3])
def test_gather_scatter_over_gloo(world):
    rng = np.random.default_rng(5)
    lens = rng.integers(30, 400, size=41)
    offsets = np.zeros(len(lens) + 1, np.int64)
    offsets[1:] = np.cumsum(lens)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, offsets, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n = len(lens)
    ii, jj = np.triu_indices(n, 1)
    for f, name in enumerate(["score", "rmsd", "tm"]):
        want = np.zeros((n, n))
        if name == "tm":
            np.fill_diagonal(want, 1.0)
        want[ii, jj] = _fake(ii, jj, f)
        want[jj, ii] = _fake(ii, jj, f)
        for rank, s, r, t in results:
            got = dict(score=s, rmsd=r, tm=t)[name]
            assert np.array_equal(got, want), (name, rank)


def test_shards_partition_and_balance():
    rng = np.random.default_rng(4)
    lens = rng.integers(50, 1001, size=300)              # BASELINE config 4 shape (mixed lengths), smaller N
    offsets = np.zeros(len(lens) + 1, np.int64)
    offsets[1:] = np.cumsum(lens)
    n = len(lens)
    for world in (1, 2, 4, 8):
        seen = np.zeros((n, n), np.int32)
        cost = []
        for r in range(world):
            pi, pj = engine.plan_shard(offsets, r, world)
            assert np.all(pi < pj)
            seen[pi, pj] += 1
            cost.append(float(np.sum(lens[pi].astype(np.float64) * lens[pj])))
        assert np.array_equal(seen, np.triu(np.ones((n, n), np.int32), 1))     # every pair exactly once
        assert max(cost) / (sum(cost) / world) < 1.02                           # cost-balanced within 2 %
    # deterministic
    a = engine.plan_shard(offsets, 1, 4)
    b = engine.plan_shard(offsets, 1, 4)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
