#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s13_pytest.txt 2>&1
echo "pytest exit $?" >> gpurun_out/s13_pytest.txt
python tools/sweep_sched.py C3 > gpurun_out/s13_sweep_c3.txt 2>&1
python tools/sweep_sched.py C5 > gpurun_out/s13_sweep_c5.txt 2>&1
tail -4 gpurun_out/s13_pytest.txt; cat gpurun_out/s13_sweep_c3.txt gpurun_out/s13_sweep_c5.txt
