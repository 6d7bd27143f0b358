"""The whole caretta-cli computation after feature extraction (align_from_structure_files, multiple_alignment.py:498-596) on one
GPU through the mirror API, step by step: pair matrix -> guide-tree distance text -> neighbor joining -> progressive alignment ->
FASTA -> RMSD / coverage / TM matrices -> their text files -> superposition.   python tools/pipeline_time.py [N] [L] [reps]"""
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from caretta_b200 import multiple_alignment as MA, neighbor_joining as NJ, synth  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    ch = synth.make_chains(n, L, 10, seed=3, family_size=20)
    rng = np.random.default_rng(0)
    letters = np.array(list("ACDEFGHIKLMNPQRSTVWY"))
    prm = dict(MA.DEFAULT_SCORE_PARAMS)
    out = None
    with tempfile.TemporaryDirectory() as td:
        for rep in range(reps):
            msa = MA.StructureMultiple.from_chains(ch)
            for p in msa.sequences:
                p.sequence = "".join(letters[rng.integers(0, 20, len(p))])
            names = [p.name for p in msa.sequences]
            t = {}
            t0 = time.perf_counter(); S = msa.make_pairwise_matrix(prm); t["pair_matrix"] = time.perf_counter() - t0
            D = S.max() - S
            t0 = time.perf_counter(); MA.write_distance_matrix(names, D, os.path.join(td, "d.txt")); t["write_guide_matrix"] = time.perf_counter() - t0
            t0 = time.perf_counter(); tree, bl = NJ.neighbor_joining(D); t["neighbor_joining"] = time.perf_counter() - t0
            t0 = time.perf_counter()
            aln = msa.progressive_align(tree, 1.0, 0.01, 1.0, 1.0, prm, dict(flexible=False))
            t["progressive_align"] = time.perf_counter() - t0
            msa.alignment = aln
            t0 = time.perf_counter(); msa.write_alignment(os.path.join(td, "a.fasta")); t["write_fasta"] = time.perf_counter() - t0
            t0 = time.perf_counter(); r, c, tm = MA.make_rmsd_coverage_tm_matrix(aln, msa.sequences, superpose_first=False); t["rmsd_cov_tm"] = time.perf_counter() - t0
            t0 = time.perf_counter()
            for nm, M in (("rmsd", r), ("coverage", c), ("tm", tm)):
                MA.write_distance_matrix(names, M, os.path.join(td, nm + ".txt"))
            t["write_3_matrices"] = time.perf_counter() - t0
            t0 = time.perf_counter(); MA.superpose(aln, msa.sequences); t["superpose"] = time.perf_counter() - t0
            out = {"N": n, "L": L, "alignment_length": len(next(iter(aln.values()))), "total_ms": 1e3 * sum(t.values()),
                   **{k + "_ms": 1e3 * v for k, v in t.items()}}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
