"""Timing of the alignment consumers on one GPU, with the CPU restatement timed beside it on a bounded
sample.  python tools/consumers_time.py [N] [A]   (defaults 5000 x 600: the target configuration's alignment)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from caretta_b200 import engine, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402


def wall(fn, reps=3):
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3, out


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
    A = int(sys.argv[2]) if len(sys.argv) > 2 else 600
    rng = np.random.default_rng(1)
    keep = rng.random((n, A)) < 0.5
    keep[:, :A // 2 + 8] = rng.random((n, A // 2 + 8)) < 0.999
    keep[:, :A // 2 + 8] |= True                                    # gap-free half: superpose() takes the core branch
    aln = -np.ones((n, A), np.int64)
    lengths = keep.sum(axis=1)
    for p in range(n):
        aln[p, keep[p]] = np.arange(lengths[p])
    ch = synth.make_chains(n, list(lengths), 10, seed=2, family_size=20)
    e = engine.Engine()
    e.set_chains(ch.coords, ch.tensors, ch.offsets)
    res = {"N": n, "A": A, "residues": int(len(ch.coords))}
    ms, _ = wall(lambda: e.coverage_gap_matrix(aln))
    res["coverage_gap_matrix"] = {"wall_ms": ms, "device_ms": e.last_elapsed_ms(), "out_bytes": n * n * 12}
    for mode, key in ((engine.SUP_CORE, "superpose_core"), (engine.SUP_REFERENCE, "superpose_reference")):
        ms, _ = wall(lambda: e.superpose(aln, mode))
        res[key] = {"wall_ms": ms, "device_ms": e.last_elapsed_ms(), "launches": e.last_launches()}
    M = rng.random((n, n)) * 100
    names = [f"protein_{i:05d}" for i in range(n)]
    ms, txt = wall(lambda: e.format_matrix(names, M))
    dev = e.last_elapsed_ms()
    res["format_matrix"] = {"wall_ms": ms, "device_ms": dev, "text_bytes": len(txt),
                            "device_GBps": (2 * M.nbytes + len(txt)) / dev / 1e6}
    seqs = ["".join(np.array(list("ACDEFGHIKLMNPQRSTVWY"))[rng.integers(0, 20, int(L))]) for L in lengths]
    ms, fa = wall(lambda: e.format_fasta(names, seqs, aln))
    res["format_fasta"] = {"wall_ms": ms, "device_ms": e.last_elapsed_ms(), "text_bytes": len(fa)}
    shp = [rng.integers(0, 1024, int(L)) for L in lengths]
    ms, cnt = wall(lambda: e.count_matrix(shp, 1024))
    res["count_matrix"] = {"wall_ms": ms, "device_ms": e.last_elapsed_ms()}
    ms, _ = wall(lambda: e.braycurtis(cnt, cnt))
    res["braycurtis"] = {"wall_ms": ms, "device_ms": e.last_elapsed_ms(), "pair_dims": n * n * 1024}
    kk = min(n, 400)
    t0 = time.perf_counter(); O.braycurtis(cnt[:kk], cnt[:kk]); res["cpu_braycurtis_ms_scaled"] = (time.perf_counter() - t0) * 1e3 * (n / kk) ** 2
    res["cpu_threads"] = O.num_threads()
    # CPU restatement beside it (bounded samples)
    k = min(n, 600)
    t0 = time.perf_counter(); O.coverage_gap_matrix(aln[:k]); res["cpu_coverage_gap_matrix_ms_scaled"] = (time.perf_counter() - t0) * 1e3 * (n / k) ** 2
    t0 = time.perf_counter(); O.format_matrix(names[:k], M[:k]); res["cpu_format_matrix_ms_scaled"] = (time.perf_counter() - t0) * 1e3 * n / k
    t0 = time.perf_counter(); "".join(f"{x:.4f}" for x in M[0]); res["python_format_matrix_ms_scaled"] = (time.perf_counter() - t0) * 1e3 * n
    coords = [ch.chain(p)[1] for p in range(k)]
    t0 = time.perf_counter(); O.superpose_reference(aln[:k], coords, 0); res["cpu_superpose_reference_ms_scaled"] = (time.perf_counter() - t0) * 1e3 * n / k
    print(json.dumps(res))


if __name__ == "__main__":
    main()
