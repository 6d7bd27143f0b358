#!/bin/bash
timeout 280 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/s58_pytest.txt
cat gpurun_out/s58_pytest.txt
if grep -q "failed\|error\|Timeout" gpurun_out/s58_pytest.txt; then exit 1; fi
for rep in 1 2; do timeout 150 python tools/msa_time.py 1000 300 2>&1 | tail -1 | tee -a gpurun_out/s58_ab.txt; done
CARETTA_B200_LEVEL_S1=0 timeout 150 python tools/msa_time.py 1000 300 2>&1 | tail -1 | sed 's/^/S1=0 /' | tee -a gpurun_out/s58_ab.txt
CARETTA_B200_TIMELINE=1 timeout 150 python tools/msa_time.py 1000 300 2>&1 | grep "msa level" | tail -16 | tee -a gpurun_out/s58_ab.txt
