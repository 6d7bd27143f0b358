"""Multi-GPU all-vs-all, one process per GPU (torchrun): pairs sharded by cost with no data-path collective, then ONE
all-gather of the packed score | rmsd | tm vectors (NCCL over NVLink in production; the same code runs over gloo with CPU
tensors in the tests).  The shard enumeration is deterministic (crt_plan_shard_pairs), so the gathered block can be scattered
into the dense [N,N] matrices the reference's consumers expect; that scatter and the one device-to-host copy happen on the
ranks that ask for the result (rank 0 by default), on the device, through the C ABI (crt_scatter_gathered).

The packed vectors are float32 in the production mode and float64 in the parity mode (CRT_FP64), so that a multi-rank float64
run is bitwise the single-rank one.  The single-process layout (all GPUs behind one call, NCCL inside the library) is
engine.MultiEngine / crt_multi_*.
"""
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
    """Host-side restatement of crt_scatter_gathered for CPU tensors (the gloo tests of the exchange logic): gathered
    [world, 3 * pad] -> dense symmetric float64 matrices with the reference's diagonals (score 0: multiple_alignment.py:161-170;
    rmsd 0, tm 1: :1019-1024)."""
    g = gathered.detach().cpu().numpy().reshape(len(shards), len(FIELDS), pad)
    out = {}
    for f, name in enumerate(FIELDS):
        M = np.zeros((n, n))
        for r, (pi, pj) in enumerate(shards):
            v = g[r, f, :len(pi)].astype(np.float64)
            M[pi, pj] = v
            M[pj, pi] = v
        if name == "tm":
            np.fill_diagonal(M, 1.0)
        out[name] = M
    return out


_SHARD_CACHE: dict = {}


def _pad_cached(offsets, world: int) -> int:
    key = (np.asarray(offsets).tobytes(), world)
    hit = _SHARD_CACHE.get(key)
    if hit is None:
        _SHARD_CACHE.clear()
        lib = _engine.load_library()
        off = np.ascontiguousarray(offsets, dtype=np.int64)
        hit = _SHARD_CACHE[key] = max(1, max(int(lib.crt_plan_shard_size(_engine._p(off), len(off) - 1, r, world)) for r in range(world)))
    return hit


def all_vs_all(eng: "_engine.Engine", prm, rank: int, world: int, group=None, out=None,
               result_ranks: Optional[Sequence[int]] = (0,)) -> Optional[Dict[str, np.ndarray]]:
    """One rank's part of the multi-GPU make_pairwise_matrix: compute the shard, pack it on the device, all-gather, and -- on
    the ranks listed in ``result_ranks`` (None = every rank) -- scatter into the dense matrices on the device and copy them to
    the host.  ``out``: optional (score, rmsd, tm) C-contiguous float64 [N,N] arrays (e.g. pinned) that receive the result.
    Returns {"score", "rmsd", "tm"} on the result ranks, None elsewhere."""
    offsets = eng._offsets
    pad = _pad_cached(offsets, world)
    f64 = int(prm.precision) == _engine.FP64
    eng.pairwise_shard(prm, rank, world)
    dev = torch.device("cuda", torch.cuda.current_device())
    local = torch.empty(len(FIELDS) * pad, dtype=torch.float64 if f64 else torch.float32, device=dev)
    eng.pack_results(local.data_ptr(), pad, f64)
    gathered = gather_packed(local, world, group)
    if result_ranks is not None and rank not in result_ranks:
        return None
    torch.cuda.current_stream().synchronize()            # the all-gather ran on torch's stream, the scatter runs on the engine's
    S, R, T = eng.scatter_gathered(gathered.data_ptr(), world, pad, f64, want_rmsd_tm=True, out=out)
    return {"score": S, "rmsd": R, "tm": T}
