#!/bin/bash
# ncu source-level stall sampling of the stage-1 v2 kernel at 1 and 2 warps per SMSP, and of fill2 at 8
mkdir -p gpurun_out
cd tools
for k in 4 8; do
ncu --set full --import-source on --clock-control none -k regex:k_fill1_v2 -s 1 -c 1 -f -o ../gpurun_out/s3_v2_k$k ./fill_probe_v2 300 10 v2 $k > ../gpurun_out/s3_ncu_v2_k$k.log 2>&1
done
ncu --set full --import-source on --clock-control none -k regex:k_fill2 -s 1 -c 1 -f -o ../gpurun_out/s3_f2_k32 ./fill_probe_v2 300 10 f2 32 > ../gpurun_out/s3_ncu_f2.log 2>&1
ls -la ../gpurun_out
