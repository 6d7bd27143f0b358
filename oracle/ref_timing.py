"""TEST / BENCH INFRASTRUCTURE ONLY.  Times the UNMODIFIED reference (numba) on the pair path and returns its scores, for
bench.py's CPU legs and parity gate.  The reference is imported from /root/reference (build container) or from the copy
oracle/build_ref.py left under oracle/_ref (GPU box).

The reference's loop (multiple_alignment.py:158-170) is a serial Python `for`; NUMBA_NUM_THREADS (bin/caretta-cli:106,
README.md:56-61) only feeds two prange functions of the fast mode, so it does nothing for this path.  Two
numbers are therefore measured:
  * as shipped: MultipleAlignment.make_pairwise_matrix itself on a small subset of chains, one core;
  * N cores: the statement inside that loop -- dtw.smith_waterman_score(arange, arange, seq_i.score_function(seq_j, **params)) --
    on a sample of pairs, fanned out over forked worker processes after the JIT warm-up.
"""
import multiprocessing as mp
import os
import time

import numpy as np

PARAMS = dict(flexible=False, gamma_tensor=7.0, gamma_coords=0.03, verbose=False)

_STATE = {}


def available() -> bool:
    from . import build_ref
    if not build_ref.root():
        return False
    try:
        import numba  # noqa: F401
    except Exception:
        return False
    return True


def load(threads: int = 0):
    """Imports the reference (once) with NUMBA_NUM_THREADS = threads (default: all cores) and warms the JIT
    (trigger_numba_compilation + one small pair, like bin/caretta-cli:107).  Returns (modules, seconds spent)."""
    if "mods" in _STATE:
        return _STATE["mods"], 0.0
    cores = threads or (os.cpu_count() or 1)
    os.environ.setdefault("NUMBA_NUM_THREADS", str(cores))
    os.environ.setdefault("OMP_NUM_THREADS", "1")            # README.md:56-61
    from . import ref_harness
    t = time.perf_counter()
    mods = ref_harness.load()
    ma, dtw = mods[0], mods[1]
    ma.trigger_numba_compilation()
    rng = np.random.default_rng(0)
    a = ma.Protein("a", rng.normal(size=(12, 10)), rng.normal(size=(12, 3)), "A" * 12)
    b = ma.Protein("b", rng.normal(size=(11, 10)), rng.normal(size=(11, 3)), "A" * 11)
    dtw.smith_waterman_score(np.arange(12), np.arange(11), a.score_function(b, **PARAMS))
    _STATE["mods"] = mods
    _STATE["numba_threads"] = int(os.environ["NUMBA_NUM_THREADS"])
    return mods, time.perf_counter() - t


def numba_threads() -> int:
    return _STATE.get("numba_threads", 0)


def _proteins(ma, ch, ids):
    out = {}
    for p in ids:
        t, c = ch.chain(int(p))
        out[int(p)] = ma.Protein(f"s{p}", np.ascontiguousarray(t), np.ascontiguousarray(c), "A" * len(t))
    return out


def _work(args):
    pi, pj = args
    ma, dtw = _STATE["mods"][0], _STATE["mods"][1]
    P = _STATE["P"]
    out = np.empty(len(pi))
    for q, (i, j) in enumerate(zip(pi, pj)):
        si, sj = P[int(i)], P[int(j)]
        out[q] = dtw.smith_waterman_score(np.arange(len(si)), np.arange(len(sj)), si.score_function(sj, **PARAMS))
    return out


def pairs_fanout(ch, pi, pj, nproc: int = 0):
    """Reference scores of the pairs (pi[q], pj[q]) with nproc forked workers.  Returns (scores, seconds, nproc)."""
    mods, _ = load()
    nproc = nproc or (os.cpu_count() or 1)
    pi, pj = np.asarray(pi), np.asarray(pj)
    _STATE["P"] = _proteins(mods[0], ch, np.unique(np.concatenate([pi, pj])))
    nproc = max(1, min(nproc, len(pi)))
    chunks = [(pi[k::nproc], pj[k::nproc]) for k in range(nproc)]
    t = time.perf_counter()
    if nproc == 1:
        parts = [_work(chunks[0])]
    else:
        with mp.get_context("fork").Pool(nproc) as pool:
            parts = pool.map(_work, chunks)
    dt = time.perf_counter() - t
    out = np.empty(len(pi))
    for k in range(nproc):
        out[k::nproc] = parts[k]
    return out, dt, nproc


def as_shipped(ch, n_chains: int):
    """MultipleAlignment.make_pairwise_matrix (the reference's own driver, serial) on the first n_chains chains.
    Returns (matrix, seconds)."""
    mods, _ = load()
    ma = mods[0]
    P = _proteins(ma, ch, range(n_chains))
    msa = ma.MultipleAlignment([P[p] for p in range(n_chains)])
    t = time.perf_counter()
    S = msa.make_pairwise_matrix(dict(PARAMS))
    return S, time.perf_counter() - t
