#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s17_pytest.txt 2>&1
echo "pytest exit $?" >> gpurun_out/s17_pytest.txt
python - > gpurun_out/s17_units.txt 2>&1 <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from caretta_b200 import synth, engine
e = engine.Engine()
for cfg, world in (("C3", 1), ("C3", 8), ("C5", 1), ("C5", 8), ("C2", 1), ("C4s", 1)):
    if cfg == "C4s":
        rng = np.random.default_rng(4)
        ch = synth.make_chains(1200, list(rng.integers(50, 1001, 1200)), 10, seed=4)
    else:
        ch = synth.config(cfg)
    e.set_chains(ch.coords, ch.tensors, ch.offsets)
    for rows in (0, 6144):
        os.environ["CARETTA_B200_UNIT_ROWS"] = str(rows)
        best = min((e.pairwise_shard(e.params(), 0, world), e.last_elapsed_ms())[1] for _ in range(3))
        print(f"{cfg} world={world} unit_rows={'adaptive' if rows == 0 else rows}: {best:9.2f} ms  launches {e.last_launches()}", flush=True)
PY
python bench.py --steps 5 --warmup 3 > gpurun_out/s17_bench.txt 2>&1
tail -3 gpurun_out/s17_pytest.txt; cat gpurun_out/s17_units.txt; cut -c1-1200 gpurun_out/s17_bench.txt
