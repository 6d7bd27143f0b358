#!/bin/bash
# final check of the round: GPU tests, smoke, bench, progressive-alignment timing
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/s61_pytest.txt 2>&1; echo "pytest exit $?" >> gpurun_out/s61_pytest.txt
timeout 120 python __graft_entry__.py smoke > gpurun_out/s61_smoke.txt 2>&1; echo "smoke exit $?" >> gpurun_out/s61_smoke.txt
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/s61_bench.txt 2>&1
for rep in 1 2; do timeout 100 python tools/msa_time.py 1000 300 2>&1 | tail -1 >> gpurun_out/s61_msa.txt; done
tail -2 gpurun_out/s61_pytest.txt; tail -2 gpurun_out/s61_smoke.txt; cut -c1-330 gpurun_out/s61_bench.txt | tail -1; cat gpurun_out/s61_msa.txt
