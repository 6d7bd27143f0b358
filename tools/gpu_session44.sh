#!/bin/bash
# multi-warp strip pipeline in k_dtw_fill: parity tests, then timing of the progressive alignment
timeout 200 python -m pytest tests/test_gpu_dp_batch.py tests/test_gpu_msa.py tests/test_gpu_flexible.py tests/test_gpu_sequence_api.py tests/test_pipeline_golden.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/s44_pytest.txt
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/s44_pytest.txt
cat gpurun_out/s44_pytest.txt
if grep -q "failed\|error\|Timeout" gpurun_out/s44_pytest.txt; then exit 1; fi
for n in 1000 5000; do timeout 150 python tools/msa_time.py $n 300 2>&1 | tail -1 | tee -a gpurun_out/s44_msa.txt; done
