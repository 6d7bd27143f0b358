#!/bin/bash
timeout 200 python -m pytest tests/test_gpu_dp_batch.py tests/test_gpu_msa.py tests/test_gpu_flexible.py tests/test_pipeline_golden.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/s46_pytest.txt
cat gpurun_out/s46_pytest.txt
if grep -q "failed\|error\|Timeout" gpurun_out/s46_pytest.txt; then exit 1; fi
for rep in 1 2; do timeout 150 python tools/msa_time.py 1000 300 2>&1 | tail -1 | tee -a gpurun_out/s46_ab.txt; done
CARETTA_B200_TIMELINE=1 timeout 150 python tools/msa_time.py 1000 300 2>&1 | grep "msa level" | tail -16 | tee -a gpurun_out/s46_ab.txt
