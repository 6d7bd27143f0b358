#!/bin/bash
mkdir -p gpurun_out
python tools/sweep_sched.py 1000 300 > gpurun_out/s5_sweep.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/s5_pytest.txt 2>&1
cat gpurun_out/s5_sweep.txt; tail -3 gpurun_out/s5_pytest.txt
