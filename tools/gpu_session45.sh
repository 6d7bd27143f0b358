#!/bin/bash
# A/B of the progressive alignment: HEAD library (libcaretta_b200_old.so) against the multi-warp k_dtw_fill, per-level timeline
for lib in old new; do
  if [ $lib = old ]; then export CARETTA_B200_LIB=$PWD/caretta_b200/libcaretta_b200_old.so; else unset CARETTA_B200_LIB; fi
  for rep in 1 2; do timeout 150 python tools/msa_time.py 1000 300 2>&1 | tail -1 | sed "s/^/$lib /" | tee -a gpurun_out/s45_ab.txt; done
  CARETTA_B200_TIMELINE=1 timeout 150 python tools/msa_time.py 1000 300 2>&1 | grep "msa level" | tail -16 | sed "s/^/$lib /" >> gpurun_out/s45_ab.txt
done
tail -40 gpurun_out/s45_ab.txt
