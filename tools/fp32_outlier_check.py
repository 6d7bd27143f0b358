"""For the C3 pairs whose fp32 stage-1 path differs from the fp64 one: how far from optimal is the fp32 path in exact arithmetic?
Sum of the float64 stage-1 scores S_T along the matched columns of each path (gap = 0: the path score is that sum)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from caretta_b200 import engine, synth
ch = synth.config("C3")
e = engine.Engine()
e.set_chains(ch.coords, ch.tensors, ch.offsets)
pi, pj = np.triu_indices(ch.n, 1)
S32 = e.pairwise_all(e.params(precision=engine.FP32))
S64 = e.pairwise_all(e.params(precision=engine.FP64))
rel = np.abs(S32[pi, pj] - S64[pi, pj]) / S64[pi, pj]
out = np.nonzero(rel > 1e-4)[0]
r32 = e.pairwise_list(e.params(precision=engine.FP32), pi[out], pj[out], want_paths=True)
r64 = e.pairwise_list(e.params(precision=engine.FP64), pi[out], pj[out], want_paths=True)
gaps, ndiff = [], []
for k, q in enumerate(out):
    t1, t2 = ch.chain(int(pi[q]))[0], ch.chain(int(pj[q]))[0]
    def path_score(r):
        a, b = r["aln_off"][k], r["aln_off"][k + 1]
        x, y = r["aln1"][a:b], r["aln2"][a:b]
        m = (x >= 0) & (y >= 0)
        d = t1[x[m]] - t2[y[m]]
        return float(np.exp(-7.0 * (d * d).sum(axis=1)).sum()), set(zip(x[m].tolist(), y[m].tolist()))
    s32, c32 = path_score(r32)
    s64, c64 = path_score(r64)
    gaps.append((s64 - s32) / s64)
    ndiff.append(len(c64 ^ c32))
gaps, ndiff = np.array(gaps), np.array(ndiff)
print(json.dumps({"outliers": len(out), "relative_score_deficit_of_fp32_path": {"median": float(np.median(gaps)), "p90": float(np.quantile(gaps, 0.9)),
                  "max": float(gaps.max()), "min": float(gaps.min())}, "columns_differing": {"median": float(np.median(ndiff)), "max": int(ndiff.max())}}))
