#!/bin/bash
# builds the GPU probe binaries of tools/ for sm_100a (they travel to the GPU box with the snapshot; the binaries are git-ignored)
cd "$(dirname "$0")"
ARCH="-gencode arch=compute_100a,code=sm_100a"
INC="-I../caretta_b200/csrc"
nvcc $ARCH -O3 -o microbench microbench.cu                                              # pipe rates of the systolic kernels' instructions
nvcc $ARCH -O3 -o microbench_fp64 microbench_fp64.cu
nvcc $ARCH -O3 -std=c++17 -lineinfo -DPROBE_V2 $INC -o fill_probe fill_probe.cu         # steady state of k_fill1_v3 / v4, k_fill2_v3
nvcc $ARCH -O3 -std=c++17 -lineinfo -o tc_probe tc_probe.cu                             # round-2 prototype: exponent tile on tcgen05.mma
nvcc $ARCH -O3 -std=c++17 -lineinfo $INC -o tc_fill_probe tc_fill_probe.cu              # k_fill1_tc: correctness + steady state
nvcc $ARCH -O3 -std=c++17 -lineinfo -DTC_PROFILE $INC -o tc_fill_probe_prof tc_fill_probe.cu   # the same with clock64 counters
nvcc $ARCH -O3 -std=c++17 -o cell_probe cell_probe.cu                                   # the stage-1 cell arithmetic in isolation
nvcc $ARCH -O3 -o alu_tp alu_tp.cu                                                      # ALU-pipe throughput / latency of the cell's instructions
nvcc $ARCH -O3 -o tmem_bw tmem_bw.cu                                                    # tcgen05.ld / st throughput, load-to-use latency
gcc -O2 -fopenmp -o tie_study tie_study.c -lm 2>/dev/null || true                       # CPU model of the tie detection (config C3)
