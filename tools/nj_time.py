"""Timing of neighbor joining on the device against the oracle port and (when it fits in seconds) noted reference complexity."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from caretta_b200 import engine
from oracle import oracle as O
e = engine.Engine()
rng = np.random.default_rng(0)
for n in (200, 1000, 2000, 5000):
    A = rng.random((n, n)); A = (A + A.T) / 2; np.fill_diagonal(A, 0)
    e.neighbor_joining(A[:64, :64])
    t0 = time.perf_counter(); tree, bl = e.neighbor_joining(A); t1 = time.perf_counter()
    dev = e.last_elapsed_ms()
    msg = f"N={n}: GPU wall {1e3*(t1-t0):.1f} ms (device {dev:.1f} ms)"
    if n <= 2000:
        t0 = time.perf_counter(); to, bo = O.neighbor_joining(A); t1 = time.perf_counter()
        msg += f"  oracle port (1 core, cached row sums) {1e3*(t1-t0):.1f} ms  identical {np.array_equal(tree, to) and np.array_equal(bl, bo)}"
    print(msg, flush=True)
