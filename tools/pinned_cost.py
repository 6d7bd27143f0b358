"""Cost of page-locked host allocations against pageable device-to-host copies (decides how large outputs are returned)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from caretta_b200 import engine
e = engine.Engine()
n = 5000
for rep in range(3):
    t0 = time.perf_counter(); a = engine.pinned_empty((n, n)); t1 = time.perf_counter()
    a[:] = 0; t2 = time.perf_counter()
    del a; t3 = time.perf_counter()
    b = np.empty((n, n)); t4 = time.perf_counter(); b[:] = 0; t5 = time.perf_counter()
    print(f"rep {rep}: pinned_empty(200 MB) {1e3*(t1-t0):.1f} ms, first touch {1e3*(t2-t1):.1f} ms, free {1e3*(t3-t2):.1f} ms; np.empty first touch {1e3*(t5-t4):.1f} ms")
rng = np.random.default_rng(0)
aln = np.where(rng.random((n, 450)) < 0.66, 1, -1).astype(np.int64)
for p in range(n):
    k = aln[p] > 0
    aln[p, k] = np.arange(k.sum())
from caretta_b200 import synth
L = (aln >= 0).sum(axis=1)
ch = synth.make_chains(n, list(L), 10, seed=2, family_size=20)
e.set_coords(ch.coords, ch.offsets)
for rep in range(3):
    t0 = time.perf_counter(); e.rmsd_cov_tm(aln); t1 = time.perf_counter()
    print(f"rmsd_cov_tm wall {1e3*(t1-t0):.1f} ms (device {e.last_elapsed_ms():.1f})")
