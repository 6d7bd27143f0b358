#!/bin/bash
# ncu --set full of the new k_dtw_fill (first level of a 300-chain tree: 128 problems of 300 x 300, one warp each)
MSA_TIME_COLD=1 timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_dtw_fill -c 1 -o gpurun_out/s60_dtw_fill python tools/msa_time.py 300 300 > gpurun_out/s60.log 2>&1
ncu -i gpurun_out/s60_dtw_fill.ncu-rep --page raw --csv > gpurun_out/s60_dtw_fill.raw.csv 2>/dev/null
ls -la gpurun_out/s60_dtw_fill.*
