#!/bin/bash
# One documented runner for the profiles/ evidence (replaces the per-session scripts of round 1).  Run under gpurun, one GPU:
#   gpurun --timeout 1500 -- bash tools/profile_run.sh r02
# Writes into gpurun_out/<tag>/:
#   launches.csv      every kernel launch of a short bench run with its device time (ncu gpu__time_duration.sum)
#   fill1.ncu-rep     ncu --set full of the dominant kernel (k_fill1_v4), 2 launches
#   fill2.ncu-rep     the same for k_fill2_v3
#   nj2.ncu-rep       the same for one segment of the in-place neighbor joining (k_nj2_persistent, 2000 nodes); nj_launches.csv
#   sanitizer_*.txt   compute-sanitizer memcheck / racecheck over the small pair cases (fp32 fills with their cp.async rings,
#                     k_trace, the float64 re-run), the affine DTW cases and the in-place neighbor joining on the golden cases
TAG=${1:-r02}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export CARETTA_B200_BATCHES=4
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > $OUT/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fill1_v4 -s 6 -c 2 -o $OUT/fill1 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs > $OUT/ncu_fill1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fill2_v3 -s 6 -c 2 -o $OUT/fill2 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs > $OUT/ncu_fill2.log 2>&1
unset CARETTA_B200_BATCHES
# in-place neighbor joining: one segment of the cooperative kernel at 2000 nodes (full set), and its launch list at 5000
ncu --set full --clock-control none --import-source on -k regex:k_nj2_persistent -c 1 -o $OUT/nj2 python tools/nj_one.py 2000 > $OUT/ncu_nj2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/nj_launches.csv python tools/nj_one.py 5000 > $OUT/nj_launches.log 2>&1
for tool in memcheck racecheck; do
    compute-sanitizer --tool $tool --print-limit 5 python -m pytest tests/test_gpu_pairs.py -q -m gpu \
        -k "small or ragged or zero_region or short_chains or c1_test" > $OUT/sanitizer_${tool}_pairs.txt 2>&1
    compute-sanitizer --tool $tool --print-limit 5 python -m pytest tests/test_gpu_dp_batch.py -q -m gpu > $OUT/sanitizer_${tool}_dp.txt 2>&1
    CARETTA_B200_NJ_INPLACE_MIN=4 compute-sanitizer --tool $tool --print-limit 5 python -m pytest tests/test_gpu_nj.py -q -m gpu -k "golden and default" > $OUT/sanitizer_${tool}_nj.txt 2>&1
done
tail -3 $OUT/sanitizer_*.txt
