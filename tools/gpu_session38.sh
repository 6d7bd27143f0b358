#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 400 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/s38_bench8.txt 2>&1
timeout 400 $TR --nproc-per-node 8 --master-port 29522 tools/run_config.py T --sample 200 --reps 3 > gpurun_out/s38_T8.txt 2>&1
for f in s38_bench8 s38_T8; do echo "== $f"; grep '^{' gpurun_out/$f.txt | cut -c1-1200; done
