#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s4_pytest.txt 2>&1
echo "pytest exit $?" >> gpurun_out/s4_pytest.txt
python tools/gpu_time.py 1000 300 > gpurun_out/s4_gpu_time.txt 2>&1
python tools/run_config.py C3 --sample 400 > gpurun_out/s4_c3.txt 2>&1
tail -5 gpurun_out/s4_pytest.txt; cat gpurun_out/s4_gpu_time.txt gpurun_out/s4_c3.txt
