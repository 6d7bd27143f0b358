"""cProfile of the steps after the progressive alignment (RMSD / coverage / TM matrices, their text files, superposition) through the
mirror API at N x L.   python tools/consumers_host_profile.py [N] [L]"""
import cProfile
import os
import pstats
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from caretta_b200 import multiple_alignment as MA, neighbor_joining as NJ, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 300
ch = synth.make_chains(n, L, 10, seed=3, family_size=20)
msa = MA.StructureMultiple.from_chains(ch)
prm = dict(MA.DEFAULT_SCORE_PARAMS)
S = msa.make_pairwise_matrix(prm)
tree, _ = NJ.neighbor_joining(S.max() - S)
aln = msa.progressive_align(tree, 1.0, 0.01, 1.0, 1.0, prm, dict(flexible=False))
msa.alignment = aln
names = [p.name for p in msa.sequences]
with tempfile.TemporaryDirectory() as td:
    def tail():
        r, c, tm = MA.make_rmsd_coverage_tm_matrix(aln, msa.sequences, superpose_first=False)
        for nm, M in (("rmsd", r), ("coverage", c), ("tm", tm)):
            MA.write_distance_matrix(names, M, os.path.join(td, nm + ".txt"))
        MA.superpose(aln, msa.sequences)
    tail(); tail()
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(3):
        tail()
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(22)
