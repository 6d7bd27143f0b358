#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 108 -c 72 --csv --log-file gpurun_out/s11_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s11_launch_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_fill|k_trace" -s 3 -c 3 -f -o gpurun_out/s11_full python tools/prof_run.py 1000 300 fp32 > gpurun_out/s11_full.log 2>&1
CARETTA_B200_TIMELINE=1 python tools/run_config.py C5 --sample 20 --reps 1 > gpurun_out/s11_c5.txt 2>&1
CARETTA_B200_TIMELINE=1 python tools/run_config.py C4 --sample 40 --reps 1 > gpurun_out/s11_c4.txt 2>&1
tail -3 gpurun_out/s11_launch_bench.log | cut -c1-300; tail -3 gpurun_out/s11_full.log; tail -30 gpurun_out/s11_c5.txt | cut -c1-400; tail -40 gpurun_out/s11_c4.txt | cut -c1-400
