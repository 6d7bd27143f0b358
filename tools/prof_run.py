import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from caretta_b200 import synth, engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
L = int(sys.argv[2]) if len(sys.argv) > 2 else 300
prec = engine.FP32 if (len(sys.argv) <= 3 or sys.argv[3] == "fp32") else engine.FP64
e = engine.Engine(); ch = synth.make_chains(n, L, 10, seed=3); e.set_chains(ch.coords, ch.tensors, ch.offsets)
e.pairwise_shard(e.params(precision=prec), 0, 1)
print("elapsed ms", e.last_elapsed_ms(), "pairs", n * (n - 1) // 2)
