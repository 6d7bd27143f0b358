#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_consumers.py -q > gpurun_out/s19_pytest.txt 2>&1
echo "pytest exit $?" >> gpurun_out/s19_pytest.txt
timeout 600 python tools/consumers_time.py 5000 600 > gpurun_out/s19_consumers.txt 2>&1
timeout 300 python tools/consumers_time.py 1000 400 >> gpurun_out/s19_consumers.txt 2>&1
tail -30 gpurun_out/s19_pytest.txt; cat gpurun_out/s19_consumers.txt
