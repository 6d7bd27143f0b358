#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nj.py tests/test_gpu_msa.py -q > gpurun_out/s25_pytest.txt 2>&1
echo "pytest exit $?" >> gpurun_out/s25_pytest.txt
timeout 600 python tools/nj_time.py > gpurun_out/s25_nj.txt 2>&1
python - > gpurun_out/s25_rmsd.txt 2>&1 <<'PY'
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
from caretta_b200 import engine, synth
for n, A in ((1000, 450), (5000, 450)):
    rng = np.random.default_rng(1)
    keep = rng.random((n, A)) < 0.66
    keep[:, :40] = True
    aln = -np.ones((n, A), np.int64); L = keep.sum(axis=1)
    for p in range(n): aln[p, keep[p]] = np.arange(L[p])
    ch = synth.make_chains(n, list(L), 10, seed=2, family_size=20)
    e = engine.Engine(); e.set_chains(ch.coords, ch.tensors, ch.offsets)
    e.rmsd_cov_tm(aln)
    t0 = time.perf_counter(); e.rmsd_cov_tm(aln); t1 = time.perf_counter()
    print(f"rmsd_cov_tm N={n} A={A}: wall {1e3*(t1-t0):.1f} ms device {e.last_elapsed_ms():.2f} ms")
    e.close()
PY
tail -5 gpurun_out/s25_pytest.txt; cat gpurun_out/s25_nj.txt gpurun_out/s25_rmsd.txt
