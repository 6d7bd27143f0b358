import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from caretta_b200 import synth, engine
e = engine.Engine(); ch = synth.config("C3")
pc, pt, po = engine.pinned_like(ch.coords), engine.pinned_like(ch.tensors), engine.pinned_like(ch.offsets)
prm = e.params()
n = ch.n
dense = tuple(engine.pinned_empty((n, n)) for _ in range(3))
for rep in range(4):
    t0 = time.perf_counter(); e.set_chains(pc, pt, po); t1 = time.perf_counter()
    e.pairwise_shard(prm, 0, 1); t2 = time.perf_counter()
    r = e.fetch(499500); t3 = time.perf_counter()
    print(f"packed: set_chains {1e3*(t1-t0):.2f} ms  shard {1e3*(t2-t1):.2f} ms (device {e.last_elapsed_ms():.2f})  fetch(pageable) {1e3*(t3-t2):.2f} ms")
for rep in range(4):
    t0 = time.perf_counter(); e.set_chains(pc, pt, po); t1 = time.perf_counter()
    e.pairwise_all(prm, want_rmsd_tm=True, out=dense); t2 = time.perf_counter()
    print(f"dense : set_chains {1e3*(t1-t0):.2f} ms  pairwise_all(3 pinned matrices) {1e3*(t2-t1):.2f} ms (device {e.last_elapsed_ms():.2f})  total {1e3*(t2-t0):.2f}")
S = e.pairwise_all(prm)
print("pageable dense check:", np.array_equal(S, dense[0]), np.array_equal(S, S.T), float(np.abs(np.diag(dense[2]) - 1).max()))
