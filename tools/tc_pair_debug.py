"""Debug of the tensor-core stage 1: all pairs (i, j) of one column chain j of a config, fp32 (no float64 re-run) against fp64
paths; prints the pairs whose path differs, whether they carry CRT_ST_TIE, and where the walks part.
   python tools/tc_pair_debug.py C3 759 [735]"""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from caretta_b200 import engine, synth

ch = synth.config(sys.argv[1])
j = int(sys.argv[2])
focus = [int(x) for x in sys.argv[3:]]
pi = np.arange(j, dtype=np.int32); pj = np.full(j, j, np.int32)
eng = engine.Engine(0)
eng.set_chains(ch.coords, ch.tensors, ch.offsets)
r64 = eng.pairwise_list(eng.params(precision=engine.FP64), pi, pj, want_paths=True)
os.environ["CARETTA_B200_TIE_RERUN"] = "0"
for tc in ("1", "0"):
    os.environ["CARETTA_B200_TC"] = tc
    r32 = eng.pairwise_list(eng.params(precision=engine.FP32), pi, pj, want_paths=True)
    ndiff = nmiss = 0
    for q in range(j):
        a32 = list(zip(r32["aln1"][r32["aln_off"][q]:r32["aln_off"][q + 1]], r32["aln2"][r32["aln_off"][q]:r32["aln_off"][q + 1]]))
        a64 = list(zip(r64["aln1"][r64["aln_off"][q]:r64["aln_off"][q + 1]], r64["aln2"][r64["aln_off"][q]:r64["aln_off"][q + 1]]))
        tie = bool(int(r32["status"][q]) & 8)
        if a32 != a64:
            ndiff += 1
            if not tie: nmiss += 1
        if (a32 != a64 and not tie) or q in focus:
            k = 0
            while k < min(len(a32), len(a64)) and a32[-1 - k] == a64[-1 - k]:
                k += 1
            print(json.dumps(dict(tc=tc, pair=[q, j], status32=int(r32["status"][q]), tie=tie, score32=float(r32["score"][q]), score64=float(r64["score"][q]),
                                  len32=len(a32), len64=len(a64), same_from_end=k,
                                  div32=[[int(v) for v in x] for x in a32[max(0, len(a32) - k - 4):len(a32) - k + 2]],
                                  div64=[[int(v) for v in x] for x in a64[max(0, len(a64) - k - 4):len(a64) - k + 2]])), flush=True)
    print(json.dumps(dict(tc=tc, column=j, pairs=j, paths_differ=ndiff, differ_and_not_marked=nmiss)), flush=True)
