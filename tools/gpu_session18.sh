#!/bin/bash
# re-entry validation: GPU tests, bench N=1, reference arm, ncu launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s18_pytest.txt 2>&1
echo "pytest exit $?" >> gpurun_out/s18_pytest.txt
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/s18_bench.txt 2>&1
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s18_bench_ref.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s18_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/s18_ncu_bench.log 2>&1
tail -3 gpurun_out/s18_pytest.txt; cut -c1-1500 gpurun_out/s18_bench.txt; cut -c1-600 gpurun_out/s18_bench_ref.txt
