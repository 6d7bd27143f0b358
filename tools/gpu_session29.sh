#!/bin/bash
# ncu --set full of the kernels behind the section-8(f) rows
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on"
timeout 600 $N -k regex:'k_nj_rebuild|k_nj_argmin' -s 6 -c 2 -f -o gpurun_out/s29_nj python tools/nj_one.py 5000 > gpurun_out/s29_nj.log 2>&1
timeout 600 $N -k regex:'k_fmt_write|k_fmt_rowlen|k_coverage_gap|k_superpose|k_rmsd_cov_tm' -c 5 -f -o gpurun_out/s29_cons python tools/consumers_time.py 5000 600 > gpurun_out/s29_cons.log 2>&1
timeout 600 $N -k regex:'k_dtw_fill|k_level_score|k_dtw_trace' -c 3 -f -o gpurun_out/s29_msa python tools/msa_time.py 1000 300 > gpurun_out/s29_msa.log 2>&1
ls -la gpurun_out/*.ncu-rep
