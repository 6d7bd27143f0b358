"""Wall time of crt_neighbor_joining at N = 5000 over repeated calls, before and after an all-vs-all run in the same process."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from caretta_b200 import engine, synth
n = 5000
rng = np.random.default_rng(0)
A = rng.random((n, n)); A = (A + A.T) / 2; np.fill_diagonal(A, 0)
e = engine.Engine()
for k in range(3):
    t0 = time.perf_counter(); e.neighbor_joining(A); print(f"fresh   call {k}: wall {1e3*(time.perf_counter()-t0):7.1f} ms device {e.last_elapsed_ms():6.1f}")
ch = synth.make_chains(n, 300, 10, seed=3, family_size=20)
e.set_chains(ch.coords, ch.tensors, ch.offsets)
S = e.pairwise_all(e.params())
D = S.max() - S
for k in range(3):
    t0 = time.perf_counter(); e.neighbor_joining(D); print(f"after pairs {k}: wall {1e3*(time.perf_counter()-t0):7.1f} ms device {e.last_elapsed_ms():6.1f}")
