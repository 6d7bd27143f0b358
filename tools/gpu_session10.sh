#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s10_pytest.txt 2>&1
echo "pytest exit $?" >> gpurun_out/s10_pytest.txt
python tools/e2e_breakdown.py > gpurun_out/s10_e2e.txt 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/s10_bench.txt 2>&1
tail -4 gpurun_out/s10_pytest.txt; cat gpurun_out/s10_e2e.txt gpurun_out/s10_bench.txt
