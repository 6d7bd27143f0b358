#!/bin/bash
mkdir -p gpurun_out
CARETTA_B200_TIMELINE=1 python - > gpurun_out/s7_timeline.txt 2>&1 <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
from caretta_b200 import synth, engine
e = engine.Engine(); ch = synth.make_chains(1000, 300, 10, seed=3); e.set_chains(ch.coords, ch.tensors, ch.offsets)
for rep in range(2):
    print("--- rep", rep, flush=True); sys.stderr.flush()
    e.pairwise_shard(e.params(), 0, 1)
    print("elapsed", e.last_elapsed_ms(), flush=True)
PY
cat gpurun_out/s7_timeline.txt
