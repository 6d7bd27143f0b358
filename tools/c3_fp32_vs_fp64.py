"""fp32 production mode against the fp64 parity mode on every pair of a configuration (default C3)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from caretta_b200 import engine, synth
cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
ch = synth.config(cfg)
e = engine.Engine()
e.set_chains(ch.coords, ch.tensors, ch.offsets)
pi, pj = np.triu_indices(ch.n, 1)
S32, R32, T32 = e.pairwise_all(e.params(precision=engine.FP32), want_rmsd_tm=True)
ms32 = e.last_elapsed_ms()
S64, R64, T64 = e.pairwise_all(e.params(precision=engine.FP64), want_rmsd_tm=True)
ms64 = e.last_elapsed_ms()
s32, s64 = S32[pi, pj], S64[pi, pj]
rel = np.abs(s32 - s64) / np.maximum(s64, 1e-300)
bad = rel > 1e-4
fam = (pi // 20) == (pj // 20)
out = {"config": cfg, "pairs": len(pi), "fp32_ms": ms32, "fp64_ms": ms64, "rel_median": float(np.median(rel)), "rel_p999": float(np.quantile(rel, 0.999)),
       "rel_max": float(rel.max()), "n_over_1e-4": int(bad.sum()), "frac_within_1e-4": float(1 - bad.mean()),
       "n_over_same_family": int((bad & fam).sum()), "same_family_pairs": int(fam.sum()),
       "score_of_outliers_median": float(np.median(s64[bad])) if bad.any() else None, "score_median_all": float(np.median(s64)),
       "score_median_same_family": float(np.median(s64[fam])),
       "rmsd_close_frac": float(np.isclose(R32[pi, pj], R64[pi, pj], rtol=1e-4, atol=1e-6).mean()),
       "tm_close_frac": float(np.isclose(T32[pi, pj], T64[pi, pj], rtol=1e-4, atol=1e-9).mean())}
print(json.dumps(out))
if bad.any():
    idx = np.nonzero(bad)[0][:8]
    r32 = e.pairwise_list(e.params(precision=engine.FP32), pi[idx], pj[idx], want_paths=True)
    r64 = e.pairwise_list(e.params(precision=engine.FP64), pi[idx], pj[idx], want_paths=True)
    for k, q in enumerate(idx):
        a, b = r32["aln_off"][k], r32["aln_off"][k + 1]
        c, d = r64["aln_off"][k], r64["aln_off"][k + 1]
        print(int(pi[q]), int(pj[q]), "s32", s32[q], "s64", s64[q], "len32", int(b - a), "len64", int(d - c), "ncommon", int(r32["ncommon"][k]), int(r64["ncommon"][k]))
# column identity on a sample: 20000 random pairs + every outlier
rng = np.random.default_rng(1)
samp = np.unique(np.concatenate([rng.choice(len(pi), 20000, replace=False), np.nonzero(bad)[0]]))
r32 = e.pairwise_list(e.params(precision=engine.FP32), pi[samp], pj[samp], want_paths=True)
r64 = e.pairwise_list(e.params(precision=engine.FP64), pi[samp], pj[samp], want_paths=True)
tot = same = 0
same_path = np.zeros(len(samp), bool)
for k in range(len(samp)):
    a, b = r32["aln_off"][k], r32["aln_off"][k + 1]
    c, d = r64["aln_off"][k], r64["aln_off"][k + 1]
    x1, y1, x2, y2 = r32["aln1"][a:b], r32["aln2"][a:b], r64["aln1"][c:d], r64["aln2"][c:d]
    c32 = set(zip(x1[(x1 >= 0) & (y1 >= 0)].tolist(), y1[(x1 >= 0) & (y1 >= 0)].tolist()))
    c64 = set(zip(x2[(x2 >= 0) & (y2 >= 0)].tolist(), y2[(x2 >= 0) & (y2 >= 0)].tolist()))
    tot += len(c64); same += len(c64 & c32); same_path[k] = c32 == c64
is_out = bad[samp]
relS = np.abs(r32["score"] - r64["score"]) / np.maximum(r64["score"], 1e-300)
print(json.dumps({"sample": len(samp), "identical_columns": same / tot, "identical_paths_random": float(same_path[~is_out].mean()),
                  "identical_paths_outliers": float(same_path[is_out].mean()) if is_out.any() else None,
                  "max_rel_where_same_path": float(relS[same_path].max()), "n_same_path_over_1e-4": int((relS[same_path] > 1e-4).sum()),
                  "n_diff_path_within_1e-4": int((relS[~same_path] <= 1e-4).sum()), "n_diff_path": int((~same_path).sum())}))
