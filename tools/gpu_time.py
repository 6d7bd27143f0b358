"""Timing of the device-resident all-vs-all step (C3 by default) with the stream / workspace knobs."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from caretta_b200 import synth, engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 300
e = engine.Engine()
ch = synth.make_chains(n, L, 10, seed=3)
e.set_chains(ch.coords, ch.tensors, ch.offsets)
peak, _ = e.fp32_peak()
npairs = n * (n - 1) // 2
for streams in (1, 2, 3, 4):
    os.environ["CARETTA_B200_STREAMS"] = str(streams)
    best = 1e9
    for rep in range(4):
        e.pairwise_shard(e.params(), 0, 1)
        best = min(best, e.last_elapsed_ms())
    cu = e.last_cell_updates()
    ph = e.last_phase_ms()
    print(f"streams={streams}: {best:.2f} ms  pairs/s {npairs / best * 1e3:.0f}  GCUPS {cu / best / 1e6:.1f}  frac(W=36) {(cu / 2 * 36) / (best * 1e-3) / peak:.3f}  phases {ph}")
