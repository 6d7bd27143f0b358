#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_msa.py -q > gpurun_out/s21_pytest.txt 2>&1
echo "pytest exit $?" >> gpurun_out/s21_pytest.txt
for cfg in "200 300" "1000 300" "2000 300"; do timeout 600 python tools/msa_time.py $cfg >> gpurun_out/s21_msa.txt 2>&1; done
CARETTA_B200_NODE_BATCH=0 timeout 600 python tools/msa_time.py 1000 300 >> gpurun_out/s21_msa.txt 2>&1
tail -30 gpurun_out/s21_pytest.txt; cat gpurun_out/s21_msa.txt
