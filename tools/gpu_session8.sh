#!/bin/bash
mkdir -p gpurun_out
( cd tools && ./fill_probe_v2 300 10 v2 8; ./fill_probe_v2 300 10 v3 4; ./fill_probe_v2 300 10 v3 8 ) > gpurun_out/s8_probe.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s8_pytest.txt 2>&1
echo "pytest exit $?" >> gpurun_out/s8_pytest.txt
python tools/gpu_time.py 1000 300 > gpurun_out/s8_gpu_time.txt 2>&1
python tools/run_config.py C3 --sample 300 > gpurun_out/s8_c3.txt 2>&1
cat gpurun_out/s8_probe.txt; tail -4 gpurun_out/s8_pytest.txt; cat gpurun_out/s8_gpu_time.txt gpurun_out/s8_c3.txt
