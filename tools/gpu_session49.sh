#!/bin/bash
MSA_TIME_COLD=1 timeout 250 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_dtw_fill -c 12 --csv --log-file gpurun_out/s49_new.csv python tools/msa_time.py 300 300 > gpurun_out/s49_new.log 2>&1
