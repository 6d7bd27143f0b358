#!/bin/bash
# launch list of one progressive alignment (device pool path), all kernels: python tools/msa_time.py 300 300, cold (one alignment only)
MSA_TIME_COLD=1 timeout 250 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/s57_msa_launches.csv python tools/msa_time.py 300 300 > gpurun_out/s57_msa.log 2>&1
tail -1 gpurun_out/s57_msa.log
