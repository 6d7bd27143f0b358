"""cProfile of MultipleAlignment.progressive_align (host side) at N x L.  python tools/msa_profile.py [N] [L]"""
import cProfile, io, os, pstats, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from caretta_b200 import multiple_alignment as MA, neighbor_joining as NJ, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 300
ch = synth.make_chains(n, L, 10, seed=3, family_size=20)
msa = MA.StructureMultiple.from_chains(ch)
prm = dict(MA.DEFAULT_SCORE_PARAMS)
S = msa.make_pairwise_matrix(prm)
tree, bl = NJ.neighbor_joining(S.max() - S)
msa.progressive_align(tree, 1.0, 0.01, 1.0, 0.03, prm, None)
t0 = time.perf_counter(); msa.progressive_align(tree, 1.0, 0.01, 1.0, 0.03, prm, None); print("wall ms", 1e3 * (time.perf_counter() - t0))
pr = cProfile.Profile(); pr.enable(); msa.progressive_align(tree, 1.0, 0.01, 1.0, 0.03, prm, None); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(18); print(s.getvalue()[:3500])
