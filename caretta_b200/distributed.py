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


_LAYOUT_CACHE: dict = {}


def _device_layout(shards, pad: int, device: torch.device):
    """Index tensors of the scatter on ``device`` (cached per shard layout): positions of every rank's valid entries in
    the flattened [world, pad] field, and the (i, j) they belong to."""
    key = (id(shards), pad, str(device))
    hit = _LAYOUT_CACHE.get(key)
    if hit is not None and hit[0] is shards:
        return hit[1]
    src = np.concatenate([r * pad + np.arange(len(pi), dtype=np.int64) for r, (pi, _) in enumerate(shards)])
    ii = np.concatenate([np.asarray(pi, dtype=np.int64) for pi, _ in shards])
    jj = np.concatenate([np.asarray(pj, dtype=np.int64) for _, pj in shards])
    lay = tuple(torch.from_numpy(a).to(device) for a in (src, ii, jj))
    _LAYOUT_CACHE.clear()                       # one layout at a time: it can be hundreds of MB at N = 5000
    _LAYOUT_CACHE[key] = (shards, lay)
    return lay


def scatter_to_device_matrices(gathered: torch.Tensor, shards, pad: int, n: int) -> torch.Tensor:
    """gathered [world, 3*pad] (any float dtype) -> [3, n, n] float64 on the same device: score, rmsd, tm, symmetric, with
    the reference's diagonals (score 0: multiple_alignment.py:161-170; rmsd 0, tm 1: :1019-1024)."""
    world = gathered.shape[0]
    src, ii, jj = _device_layout(shards, pad, gathered.device)
    g = gathered.view(world, len(FIELDS), pad)
    out = torch.zeros(len(FIELDS), n, n, dtype=torch.float64, device=gathered.device)
    for f, name in enumerate(FIELDS):
        v = g[:, f, :].reshape(-1).index_select(0, src).to(torch.float64)
        out[f].index_put_((ii, jj), v)
        out[f].index_put_((jj, ii), v)
        if name == "tm":
            out[f].diagonal().fill_(1.0)
    return out


def scatter_to_matrices(gathered: torch.Tensor, shards, pad: int, n: int, out: Optional[torch.Tensor] = None) -> Dict[str, np.ndarray]:
    """Dense symmetric float64 matrices on the host.  The scatter runs where ``gathered`` lives (the GPU in production);
    ``out``: optional pinned [3, n, n] float64 host tensor that receives the copy."""
    dense = scatter_to_device_matrices(gathered.detach(), shards, pad, n)
    if out is None:
        host = dense.cpu()
    else:
        out.copy_(dense)
        host = out
    h = host.numpy()
    return {name: h[f] for f, name in enumerate(FIELDS)}


_SHARD_CACHE: dict = {}


def _shard_layout_cached(offsets, world: int):
    key = (np.asarray(offsets).tobytes(), world)
    hit = _SHARD_CACHE.get(key)
    if hit is None:
        _SHARD_CACHE.clear()
        hit = _SHARD_CACHE[key] = shard_layout(offsets, world)
    return hit


def all_vs_all(eng: "_engine.Engine", prm, rank: int, world: int, group=None, out: Optional[torch.Tensor] = None) -> Dict[str, np.ndarray]:
    """Full pipeline on one rank: compute the shard on the GPU, all-gather the packed vectors, scatter them into the dense
    matrices on the GPU, copy to the host.  Returns score/rmsd/tm [N,N] float64 (views of ``out`` when given)."""
    offsets = eng._offsets
    n = len(offsets) - 1
    shards, pad = _shard_layout_cached(offsets, world)
    eng.pairwise_shard(prm, rank, world)
    dev = torch.device("cuda", torch.cuda.current_device())
    local = torch.zeros(len(FIELDS) * pad, dtype=torch.float32, device=dev)
    p = local.data_ptr()
    eng.fetch_device(p, p + 4 * pad, p + 8 * pad, pad)
    gathered = gather_packed(local, world, group)
    return scatter_to_matrices(gathered, shards, pad, n, out=out)
