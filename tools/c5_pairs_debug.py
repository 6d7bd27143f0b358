"""Debug: given pairs of a config, fp32 (no re-run) vs fp64 paths: where they diverge, whether the pair is marked under loose
thresholds.   python tools/c5_pairs_debug.py C5 239,331 147,365 243,462"""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from caretta_b200 import engine, synth

ch = synth.config(sys.argv[1])
pairs = [tuple(int(x) for x in a.split(",")) for a in sys.argv[2:]]
pi = np.array([p[0] for p in pairs], np.int32); pj = np.array([p[1] for p in pairs], np.int32)
eng = engine.Engine(0)
eng.set_chains(ch.coords, ch.tensors, ch.offsets)
r64 = eng.pairwise_list(eng.params(precision=engine.FP64), pi, pj, want_paths=True)
for c, eps in [(8, 1e-4), (1e4, 1e-4), (8, 1e-2), (1e6, 1e-1)]:
    os.environ["CARETTA_B200_TIE_C"] = str(c); os.environ["CARETTA_B200_TIE_EPS"] = str(eps); os.environ["CARETTA_B200_TIE_RERUN"] = "0"
    r32 = eng.pairwise_list(eng.params(precision=engine.FP32), pi, pj, want_paths=True)
    for q, p in enumerate(pairs):
        a32 = list(zip(r32["aln1"][r32["aln_off"][q]:r32["aln_off"][q + 1]], r32["aln2"][r32["aln_off"][q]:r32["aln_off"][q + 1]]))
        a64 = list(zip(r64["aln1"][r64["aln_off"][q]:r64["aln_off"][q + 1]], r64["aln2"][r64["aln_off"][q]:r64["aln_off"][q + 1]]))
        # first divergence from the END of the paths (the walk starts there)
        k = 0
        while k < min(len(a32), len(a64)) and a32[-1 - k] == a64[-1 - k]:
            k += 1
        c32 = set(x for x in a32 if x[0] >= 0 and x[1] >= 0); c64 = set(x for x in a64 if x[0] >= 0 and x[1] >= 0)
        print(json.dumps(dict(c=c, eps=eps, pair=p, status32=int(r32["status"][q]), score32=float(r32["score"][q]), score64=float(r64["score"][q]),
                              len32=len(a32), len64=len(a64), same_from_end=k, end32=[int(v) for v in a32[-1]], end64=[int(v) for v in a64[-1]],
                              div32=[[int(v) for v in x] for x in a32[max(0, len(a32) - k - 3):len(a32) - k + 1]],
                              div64=[[int(v) for v in x] for x in a64[max(0, len(a64) - k - 3):len(a64) - k + 1]],
                              cols_only32=len(c32 - c64), cols_only64=len(c64 - c32))), flush=True)
