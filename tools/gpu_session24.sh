#!/bin/bash
# per-kernel device time of the level-batched progressive alignment and of the consumers (ncu launch lists; shares, not absolutes)
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_level|k_dtw|k_fill|k_trace|k_chain_of|k_tensor|k_centroid|k_prep' -c 600 --csv \
    --log-file gpurun_out/s24_launches_msa.csv python tools/msa_time.py 1000 300 > gpurun_out/s24_msa.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_aln_bits|k_coverage|k_core|k_superpose|k_fmt|k_fasta|k_rmsd' -c 200 --csv \
    --log-file gpurun_out/s24_launches_consumers.csv python tools/consumers_time.py 5000 600 > gpurun_out/s24_cons.log 2>&1
python - <<'PY'
import csv, collections
for f in ("gpurun_out/s24_launches_msa.csv", "gpurun_out/s24_launches_consumers.csv"):
    rows = [r for r in csv.reader(l for l in open(f) if l.startswith('"'))]
    hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        try: v = float(r[vi].replace(",", ""))
        except ValueError: continue
        k = r[ki].split("(")[0][:60]
        agg[k][0] += 1; agg[k][1] += v
    print(f)
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"  {k:60s} launches {n:5d}  total {t/1e6:10.3f} ms")
PY
