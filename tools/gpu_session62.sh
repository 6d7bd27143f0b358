#!/bin/bash
timeout 200 python -m pytest tests/test_gpu_msa.py tests/test_gpu_flexible.py tests/test_pipeline_golden.py tests/test_gpu_sequence_api.py -m gpu -q 2>&1 | tail -3 > gpurun_out/s62_pytest.txt
cat gpurun_out/s62_pytest.txt
for rep in 1 2 3; do timeout 100 python tools/msa_time.py 1000 300 2>&1 | tail -1 >> gpurun_out/s62_msa.txt; done
cut -c150-330 gpurun_out/s62_msa.txt
