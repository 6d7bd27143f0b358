"""Multi-GPU all-vs-all: one process per GPU (torchrun), pairs sharded by cost with no data-path collective, then ONE
all-gather of the packed score / RMSD / TM vectors (NCCL over NVLink in production; the same code runs over gloo
with CPU tensors in the tests).  The shard enumeration is deterministic (crt_plan_shard_pairs), so every rank can
scatter every other rank's packed vector into the dense [N,N] matrices the reference's consumers expect."""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist

from . import engine as _engine

FIELDS = ("score", "rmsd", "tm")


def shard_layout(offsets, world: int):
    """[(pair_i, pair_j)] per rank and the padded per-rank count used for the equal-size all-gather."""
    shards = [_engine.plan_shard(offsets, r, world) for r in range(world)]
    pad = max(1, max(len(s[0]) for s in shards))
    return shards, pad


def gather_packed(local: torch.Tensor, world: int, group=None) -> torch.Tensor:
    """local: [n_fields * pad] on this rank's device -> [world, n_fields * pad] on every rank."""
    out = torch.empty(world * local.numel(), dtype=local.dtype, device=local.device)
    if world > 1:
        dist.all_gather_into_tensor(out, local, group=group)
    else:
        out.copy_(local)
    return out.view(world, local.numel())


def scatter_to_matrices(gathered: torch.Tensor, shards, pad: int, n: int) -> Dict[str, np.ndarray]:
    """gathered [world, 3*pad] -> dense symmetric float64 matrices with the reference's diagonals
    (score 0: multiple_alignment.py:161-170; rmsd 0, tm 1: :1019-1024)."""
    g = gathered.detach().to("cpu", torch.float64).numpy()
    out = {}
    for f, name in enumerate(FIELDS):
        m = np.zeros((n, n))
        if name == "tm":
            np.fill_diagonal(m, 1.0)
        for r, (pi, pj) in enumerate(shards):
            v = g[r, f * pad:f * pad + len(pi)]
            m[pi, pj] = v
            m[pj, pi] = v
        out[name] = m
    return out


def all_vs_all(eng: "_engine.Engine", prm, rank: int, world: int, group=None) -> Dict[str, np.ndarray]:
    """Full pipeline on one rank: compute the shard on the GPU, all-gather, scatter.  Returns score/rmsd/tm [N,N]."""
    offsets = eng._offsets
    n = len(offsets) - 1
    shards, pad = shard_layout(offsets, world)
    eng.pairwise_shard(prm, rank, world)
    dev = torch.device("cuda", torch.cuda.current_device())
    local = torch.zeros(len(FIELDS) * pad, dtype=torch.float32, device=dev)
    p = local.data_ptr()
    eng.fetch_device(p, p + 4 * pad, p + 8 * pad, pad)
    gathered = gather_packed(local, world, group)
    return scatter_to_matrices(gathered, shards, pad, n)
