#!/bin/bash
mkdir -p gpurun_out
( cd tools && ./fill_probe_v2 300 10 f2 32; ./fill_probe_v2 300 10 f23 4; ./fill_probe_v2 300 10 f23 8; ./fill_probe_v2 300 10 f23 16; ./fill_probe_v2 300 10 f23 32 ) 2>&1 | grep -v device > gpurun_out/s9_probe.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s9_pytest.txt 2>&1
echo "pytest exit $?" >> gpurun_out/s9_pytest.txt
bash tools/ab_time.sh 1000 300 > gpurun_out/s9_ab.txt 2>&1
python tools/run_config.py C3 --sample 300 > gpurun_out/s9_c3.txt 2>&1
cat gpurun_out/s9_probe.txt; tail -4 gpurun_out/s9_pytest.txt; cat gpurun_out/s9_ab.txt gpurun_out/s9_c3.txt
