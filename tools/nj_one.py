"""One neighbor-joining run of N random nodes (for ncu launch lists).  python tools/nj_one.py [N]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from caretta_b200 import engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
rng = np.random.default_rng(0)
A = rng.random((n, n)); A = (A + A.T) / 2; np.fill_diagonal(A, 0)
e = engine.Engine()
t0 = time.perf_counter(); e.neighbor_joining(A); t1 = time.perf_counter()
print(f"N={n}: wall {1e3 * (t1 - t0):.1f} ms, device {e.last_elapsed_ms():.1f} ms")
