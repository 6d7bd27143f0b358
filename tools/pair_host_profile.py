"""cProfile of the mirror's make_pairwise_matrix (N chains x L): where the wall time outside the device goes.
python tools/pair_host_profile.py [N] [L]"""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from caretta_b200 import multiple_alignment as MA, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 300
ch = synth.make_chains(n, L, 10, seed=3, family_size=20)
msa = MA.StructureMultiple.from_chains(ch)
prm = dict(MA.DEFAULT_SCORE_PARAMS)
for _ in range(3):
    msa.make_pairwise_matrix(prm)
ts = []
for _ in range(5):
    t0 = time.perf_counter()
    msa.make_pairwise_matrix(prm)
    ts.append((time.perf_counter() - t0) * 1e3)
print("make_pairwise_matrix wall ms:", [round(t, 2) for t in ts])
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    msa.make_pairwise_matrix(prm)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
