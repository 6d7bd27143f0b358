#!/bin/bash
timeout 200 python -m pytest tests/test_gpu_dp_batch.py tests/test_gpu_msa.py tests/test_gpu_flexible.py tests/test_pipeline_golden.py tests/test_gpu_sequence_api.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/s56_pytest.txt
cat gpurun_out/s56_pytest.txt
if grep -q "failed\|error\|Timeout" gpurun_out/s56_pytest.txt; then exit 1; fi
MSA_TIME_COLD=1 timeout 250 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_dtw_fill -c 12 --csv --log-file gpurun_out/s56_new.csv python tools/msa_time.py 300 300 > gpurun_out/s56_new.log 2>&1
for rep in 1 2; do timeout 150 python tools/msa_time.py 1000 300 2>&1 | tail -1 | tee -a gpurun_out/s56_ab.txt; done
