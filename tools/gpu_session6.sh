#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s6_pytest.txt 2>&1
echo "pytest exit $?" >> gpurun_out/s6_pytest.txt
bash tools/ab_time.sh 1000 300 > gpurun_out/s6_ab.txt 2>&1
python tools/run_config.py C3 --sample 300 > gpurun_out/s6_c3.txt 2>&1
tail -4 gpurun_out/s6_pytest.txt; cat gpurun_out/s6_ab.txt gpurun_out/s6_c3.txt
