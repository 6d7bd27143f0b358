#!/bin/bash
mkdir -p gpurun_out
( cd tools && ./microbench ) > gpurun_out/s2_microbench.txt 2>&1
( cd tools && ./fill_probe_v2 300 10 | grep -v "fill2\|fill1<10,6>" ) > gpurun_out/s2_probe_v2.txt 2>&1
( cd tools && ./fill_probe_v2s 300 10 | grep "v2" ) > gpurun_out/s2_probe_v2s.txt 2>&1
cat gpurun_out/s2_microbench.txt gpurun_out/s2_probe_v2.txt gpurun_out/s2_probe_v2s.txt
