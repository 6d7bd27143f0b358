#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/s16_bench8.txt 2>&1
$TR --nproc-per-node 8 --master-port 29512 tools/run_config.py T --sample 200 --reps 3 > gpurun_out/s16_T8.txt 2>&1
$TR --nproc-per-node 8 --master-port 29513 tools/run_config.py C4 --sample 60 --reps 2 > gpurun_out/s16_C4_8.txt 2>&1
$TR --nproc-per-node 8 --master-port 29514 tools/run_config.py C5 --sample 20 --reps 2 > gpurun_out/s16_C5_8.txt 2>&1
$TR --nproc-per-node 2 --master-port 29515 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/s16_bench2.txt 2>&1
$TR --nproc-per-node 4 --master-port 29516 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/s16_bench4.txt 2>&1
for f in s16_bench8 s16_T8 s16_C4_8 s16_C5_8 s16_bench2 s16_bench4; do echo "== $f"; grep '^{' gpurun_out/$f.txt | cut -c1-900; done
