"""Summarises an `ncu --page source --csv` export: total stall samples by reason, and the top instructions by
not-issued samples with their dominant reasons.   python tools/ncu_src_summary.py src.csv [top]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
reasons = [h for h in hdr if h.startswith("stall_") and "(Not Issued)" not in h]
tot = collections.Counter()
per = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    def f(h):
        try: return float(r[ix[h]])
        except ValueError: return 0.0
    d = {h: f(h) for h in reasons}
    for h, v in d.items(): tot[h] += v
    per.append((f("# Samples"), r[ix["Address"]], r[ix["Source"]], d, f("Instructions Executed")))
S = sum(tot.values())
print("total samples", S)
for h, v in tot.most_common():
    if v: print(f"  {h:24s} {v:9.0f}  {100 * v / S:5.1f}%")
per.sort(key=lambda x: -x[0])
print("top instructions by samples:")
for n, addr, src, d, ex in per[:top]:
    dom = sorted(d.items(), key=lambda kv: -kv[1])[:3]
    print(f"  {n:7.0f} {100 * n / S:5.1f}%  {src[:70]:70s} " + ", ".join(f"{k[6:]}={v:.0f}" for k, v in dom if v))
