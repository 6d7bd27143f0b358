#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_level|k_dtw|k_fill|k_trace|k_chain_of|k_tensor|k_centroid|k_prep|k_pool|k_rows' -c 900 --csv \
    --log-file gpurun_out/s37_launches_msa_pool.csv python tools/msa_time.py 1000 300 > gpurun_out/s37_msa.log 2>&1
python - <<'PY'
import csv, collections
f = "gpurun_out/s37_launches_msa_pool.csv"
rows = [r for r in csv.reader(l for l in open(f) if l.startswith('"'))]
hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0, []])
for r in rows[1:]:
    try: v = float(r[vi].replace(",", ""))
    except ValueError: continue
    k = r[ki].split("(")[0][:60]
    agg[k][0] += 1; agg[k][1] += v; agg[k][2].append(v)
for k, (n, t, vs) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"  {k:60s} launches {n:5d}  total {t/1e6:9.3f} ms   last {vs[-1]/1e3:8.1f} us  median {sorted(vs)[len(vs)//2]/1e3:8.1f} us")
PY
