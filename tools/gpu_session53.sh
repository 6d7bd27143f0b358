#!/bin/bash
for lib in pf1 pf2; do
  export CARETTA_B200_LIB=$PWD/caretta_b200/lib$lib.so
  MSA_TIME_COLD=1 timeout 250 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_dtw_fill -c 12 --csv --log-file gpurun_out/s53_$lib.csv python tools/msa_time.py 300 300 > gpurun_out/s53_$lib.log 2>&1
done
