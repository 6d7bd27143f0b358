#!/bin/bash
# GPU session 1: sanity (tests), pipe microbenchmarks, steady-state fill probe, phase timings.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/s1_smi.txt 2>&1
( cd tools && ./microbench ) > gpurun_out/s1_microbench.txt 2>&1
( cd tools && ./fill_probe 300 10 ) > gpurun_out/s1_probe.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s1_pytest.txt 2>&1
echo "pytest exit $?" >> gpurun_out/s1_pytest.txt
python tools/gpu_time.py 1000 300 > gpurun_out/s1_gpu_time.txt 2>&1
CARETTA_B200_WORKSPACE_MB=60000 CARETTA_B200_STREAMS=1 python - > gpurun_out/s1_onebatch.txt 2>&1 <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
from caretta_b200 import synth, engine
e = engine.Engine(); ch = synth.make_chains(1000, 300, 10, seed=3); e.set_chains(ch.coords, ch.tensors, ch.offsets)
for rep in range(3):
    e.pairwise_shard(e.params(), 0, 1)
    print("one batch, one stream:", e.last_elapsed_ms(), e.last_phase_ms(), e.last_launches())
PY
tail -3 gpurun_out/s1_pytest.txt; cat gpurun_out/s1_microbench.txt gpurun_out/s1_probe.txt gpurun_out/s1_gpu_time.txt gpurun_out/s1_onebatch.txt
