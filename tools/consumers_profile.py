"""cProfile of the consumer calls at N x L (wall-time breakdown of the host side).  python tools/consumers_profile.py [N] [L]"""
import cProfile, io, os, pstats, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from caretta_b200 import multiple_alignment as MA, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 300
A = int(L * 1.5)
rng = np.random.default_rng(1)
keep = rng.random((n, A)) < (L / A)
keep[:, :A // 2 + 4] = True
aln_arr = -np.ones((n, A), np.int64)
lengths = keep.sum(axis=1)
for p in range(n):
    aln_arr[p, keep[p]] = np.arange(lengths[p])
ch = synth.make_chains(n, list(lengths), 10, seed=2, family_size=20)
msa = MA.StructureMultiple.from_chains(ch)
aln = {p.name: aln_arr[q] for q, p in enumerate(msa.sequences)}
eng = MA.get_engine()
names = [p.name for p in msa.sequences]
M = rng.random((n, n))
for what, fn in (("superpose", lambda: MA.superpose(aln, msa.sequences)),
                 ("rmsd_cov_tm", lambda: MA.make_rmsd_coverage_tm_matrix(aln, msa.sequences, superpose_first=False)),
                 ("format_matrix", lambda: eng.format_matrix(names, M))):
    fn()
    t0 = time.perf_counter(); fn(); wall = time.perf_counter() - t0
    dev = eng.last_elapsed_ms()
    pr = cProfile.Profile(); pr.enable(); fn(); pr.disable()
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(12)
    print(f"==== {what}: wall {wall * 1e3:.1f} ms, last device ms {dev:.2f}")
    print("\n".join(s.getvalue().splitlines()[4:24]))
