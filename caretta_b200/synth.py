"""Seeded synthetic chain families.  Identical arrays go to the oracle, the
reference (golden generation) and the GPU engine.

Families of 20 chains.  A parent has ``ceil(1.1 L)`` residues: CA trace = cumulative sum of 3.8 A steps
along random unit vectors, shape tensors ~ N(0, 0.3).  A child keeps a sorted random subset of exactly L
parent residues (this produces indels between siblings), adds N(0, 0.5 A) coordinate noise and N(0, 0.02)
tensor noise, and is moved by a random proper rotation and an N(0, 20 A) translation.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

FAMILY_SIZE = 20


@dataclass
class Chains:
    """Packed residue-major chain set: chain p owns rows offsets[p]:offsets[p+1]."""
    coords: np.ndarray   # float64 [sum L, 3]
    tensors: np.ndarray  # float64 [sum L, d]
    offsets: np.ndarray  # int64   [N + 1]

    @property
    def n(self) -> int:
        return len(self.offsets) - 1

    @property
    def d(self) -> int:
        return self.tensors.shape[1]

    def length(self, p: int) -> int:
        return int(self.offsets[p + 1] - self.offsets[p])

    def chain(self, p: int):
        s, e = int(self.offsets[p]), int(self.offsets[p + 1])
        return self.tensors[s:e], self.coords[s:e]

    @property
    def lengths(self) -> np.ndarray:
        return np.diff(self.offsets)


def _random_rotation(rng: np.random.Generator) -> np.ndarray:
    q, r = np.linalg.qr(rng.normal(size=(3, 3)))
    q = q * np.sign(np.diag(r))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    return q


def make_chains(n_chains: int, length, d: int = 10, seed: int = 0,
                family_size: int = FAMILY_SIZE) -> Chains:
    """``length`` is an int (all chains equal) or a sequence of per-chain lengths."""
    rng = np.random.default_rng(seed)
    if np.isscalar(length):
        lengths = np.full(n_chains, int(length), dtype=np.int64)
    else:
        lengths = np.asarray(length, dtype=np.int64)
        assert len(lengths) == n_chains
    coords: List[np.ndarray] = []
    tensors: List[np.ndarray] = []
    p = 0
    while p < n_chains:
        fam = list(range(p, min(p + family_size, n_chains)))
        lmax = int(lengths[fam].max())
        lp = int(math.ceil(1.1 * lmax))
        steps = rng.normal(size=(lp, 3))
        steps /= np.linalg.norm(steps, axis=1, keepdims=True)
        parent_xyz = np.cumsum(3.8 * steps, axis=0)
        parent_t = rng.normal(0.0, 0.3, size=(lp, d))
        for q in fam:
            l = int(lengths[q])
            keep = np.sort(rng.choice(lp, size=l, replace=False))
            xyz = parent_xyz[keep] + rng.normal(0.0, 0.5, size=(l, 3))
            rot = _random_rotation(rng)
            xyz = xyz @ rot + rng.normal(0.0, 20.0, size=3)
            t = parent_t[keep] + rng.normal(0.0, 0.02, size=(l, d))
            coords.append(np.ascontiguousarray(xyz, dtype=np.float64))
            tensors.append(np.ascontiguousarray(t, dtype=np.float64))
        p += len(fam)
    offsets = np.zeros(n_chains + 1, dtype=np.int64)
    offsets[1:] = np.cumsum(lengths)
    return Chains(np.concatenate(coords), np.concatenate(tensors), offsets)


# The configurations of BASELINE.json / SURVEY.md §8d.
def config(name: str, d: int = 10) -> Chains:
    name = name.upper()
    if name == "C2":
        return make_chains(200, 80, d, seed=2)
    if name == "C3":
        return make_chains(1000, 300, d, seed=3)
    if name == "C4":
        rng = np.random.default_rng(4)
        return make_chains(5000, rng.integers(50, 1001, size=5000), d, seed=4)
    if name == "C5":
        return make_chains(500, 1500, d, seed=5)
    if name == "T":
        return make_chains(5000, 300, d, seed=6)
    raise ValueError(f"unknown config {name!r}")


def read_ca_coords(pdb_path) -> np.ndarray:
    """Minimal ATOM/CA reader for the reference's sample PDBs (first model, first altloc)."""
    out = []
    seen = set()
    with open(pdb_path) as fh:
        for line in fh:
            if line.startswith("ENDMDL"):
                break
            if line.startswith("ATOM") and line[12:16].strip() == "CA":
                key = (line[21], line[22:27])
                if key in seen:
                    continue
                seen.add(key)
                out.append((float(line[30:38]), float(line[38:46]), float(line[46:54])))
    return np.asarray(out, dtype=np.float64)
