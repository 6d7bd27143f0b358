#!/bin/bash
# round-end style check: GPU tests, smoke, bench (both arms), launch list of the bench command
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/s59_pytest.txt 2>&1; echo "pytest exit $?" >> gpurun_out/s59_pytest.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/s59_smoke.txt 2>&1; echo "smoke exit $?" >> gpurun_out/s59_smoke.txt
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s59_bench_ref.txt 2>&1
timeout 600 python bench.py > gpurun_out/s59_bench.txt 2>&1
tail -2 gpurun_out/s59_pytest.txt; tail -2 gpurun_out/s59_smoke.txt; cut -c1-400 gpurun_out/s59_bench.txt | tail -1; cut -c1-200 gpurun_out/s59_bench_ref.txt | tail -1
