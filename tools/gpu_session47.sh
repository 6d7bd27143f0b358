#!/bin/bash
# per-launch durations of k_dtw_fill: HEAD library against the multi-warp version
for lib in old new; do
  if [ $lib = old ]; then export CARETTA_B200_LIB=$PWD/caretta_b200/libcaretta_b200_old.so; else unset CARETTA_B200_LIB; fi
  MSA_TIME_COLD=1 timeout 250 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_dtw_fill -c 40 --csv --log-file gpurun_out/s47_$lib.csv python tools/msa_time.py 300 300 > gpurun_out/s47_$lib.log 2>&1
  grep -c k_dtw_fill gpurun_out/s47_$lib.csv
done
