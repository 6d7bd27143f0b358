#!/bin/bash
# compute-sanitizer memcheck over the entry points added at the end of round 1 (flexible scoring, Protein methods, generic driver)
CS="/usr/local/cuda/bin/compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20"
timeout 400 $CS python -m pytest tests/test_gpu_flexible.py tests/test_gpu_sequence_api.py -m gpu -q -x > gpurun_out/s43_memcheck.txt 2>&1
echo "memcheck exit $?" >> gpurun_out/s43_memcheck.txt
tail -8 gpurun_out/s43_memcheck.txt
