#!/bin/bash
mkdir -p gpurun_out
CARETTA_B200_TIMELINE=1 python tools/run_config.py C4 --sample 0 --reps 1 > gpurun_out/s14_c4.txt 2>&1
python tools/run_config.py C4 --sample 40 --reps 2 > gpurun_out/s14_c4_plain.txt 2>&1
tail -2 gpurun_out/s14_c4.txt | cut -c1-600;  cat gpurun_out/s14_c4_plain.txt
