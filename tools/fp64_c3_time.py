"""Device time of one all-vs-all step of config C3 (1000 x 300) in the float64 parity mode and in fp32.  python tools/fp64_c3_time.py [N]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from caretta_b200 import engine, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
ch = synth.make_chains(n, 300, 10, seed=3)
e = engine.Engine()
e.set_chains(ch.coords, ch.tensors, ch.offsets)
for name, prec in (("fp64", engine.FP64), ("fp32", engine.FP32)):
    best = 1e30
    for _ in range(3):
        e.pairwise_shard(e.params(7.0, 0.03, prec), 0, 1)
        best = min(best, e.last_elapsed_ms())
    print(f"{name}: {best:.2f} ms per step of {n * (n - 1) // 2} pairs")
