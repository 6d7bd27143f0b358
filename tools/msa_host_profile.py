"""cProfile of the host side of progressive_align (N chains x L residues) on one GPU: where the wall time outside the kernels goes.
python tools/msa_host_profile.py [N] [L]"""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from caretta_b200 import multiple_alignment as MA, neighbor_joining as NJ, synth  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    ch = synth.make_chains(n, L, 10, seed=3, family_size=20)
    msa = MA.StructureMultiple.from_chains(ch)
    prm = dict(MA.DEFAULT_SCORE_PARAMS)
    S = msa.make_pairwise_matrix(prm)
    tree, _ = NJ.neighbor_joining(S.max() - S)
    for _ in range(2):
        msa.progressive_align(tree, 1.0, 0.01, 1.0, 0.03, prm, dict(flexible=False))
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        msa.progressive_align(tree, 1.0, 0.01, 1.0, 0.03, prm, dict(flexible=False))
        ts.append((time.perf_counter() - t0) * 1e3)
    print("progressive_align wall ms:", [round(t, 2) for t in ts])
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(3):
        msa.progressive_align(tree, 1.0, 0.01, 1.0, 0.03, prm, dict(flexible=False))
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
    pstats.Stats(pr).sort_stats("tottime").print_stats(25)


if __name__ == "__main__":
    main()
