#!/bin/bash
# builds the GPU probe binaries of tools/ for sm_100a (they travel to the GPU box with the snapshot)
cd "$(dirname "$0")"
ARCH="-gencode arch=compute_100a,code=sm_100a"
nvcc $ARCH -O3 -o microbench microbench.cu
nvcc $ARCH -O3 -std=c++17 -lineinfo -DPROBE_V2 -I../caretta_b200/csrc -o fill_probe fill_probe.cu
