#!/bin/bash
# compute-sanitizer memcheck over the kernels added in the second half of the round (consumers, level / pool, NJ rebuild)
mkdir -p gpurun_out
CS="/usr/local/cuda/bin/compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20"
timeout 1500 $CS python -m pytest tests/test_gpu_consumers.py -q -x -k "not printf and not large" > gpurun_out/s34_cons.txt 2>&1; echo "exit $?" >> gpurun_out/s34_cons.txt
timeout 1500 $CS python -m pytest tests/test_gpu_msa.py tests/test_gpu_nj.py tests/test_pipeline_golden.py -q -x > gpurun_out/s34_msa.txt 2>&1; echo "exit $?" >> gpurun_out/s34_msa.txt
for f in s34_cons s34_msa; do echo "== $f"; grep -c "Invalid\|ERROR SUMMARY" gpurun_out/$f.txt; grep "ERROR SUMMARY\|passed\|failed\|exit" gpurun_out/$f.txt | tail -4; grep -m5 -A8 "Invalid" gpurun_out/$f.txt; done
