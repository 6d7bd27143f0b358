"""fp32 production mode against the float64 parity mode on EVERY pair of a config, for several settings of the tie detection
(crt_fill1_v4.cuh): how many pairs are marked, how long their float64 re-run takes, how many pairs still miss 1e-4, and -- with
the re-run switched off -- how many of the pairs that differ carry no mark (what the detection misses).

    python tools/tie_gpu_study.py C3 [C2 C5 ...]
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from caretta_b200 import engine, synth  # noqa: E402


def run(eng, ch, prec):
    prm = eng.params(precision=prec)
    eng.pairwise_shard(prm, 0, 1)
    eng.pairwise_shard(prm, 0, 1)
    ms = eng.last_elapsed_ms()
    n = eng.shard_size(0, 1)
    r = eng.fetch(n)
    return r, ms


def main():
    eng = engine.Engine(0)
    for name in sys.argv[1:] or ["C3"]:
        ch = synth.config(name)
        eng.set_chains(ch.coords, ch.tensors, ch.offsets)
        r64, ms64 = run(eng, ch, engine.FP64)
        s64 = r64["score"]
        pi, pj = eng.shard_pairs(0, 1)
        grid = [(c, 1e-4, 0) for c in (1, 2, 4, 8, 16, 64)] + [(8, 1e-4, 1), (4, 3e-4, 0), (4, 3e-5, 0)]
        if os.environ.get('TIE_GRID'):
            grid = json.loads(os.environ['TIE_GRID'])
        for c, eps, rerun in grid:
            os.environ["CARETTA_B200_TIE_C"] = str(c)
            os.environ["CARETTA_B200_TIE_EPS"] = str(eps)
            os.environ["CARETTA_B200_TIE_RERUN"] = str(rerun)
            r32, ms32 = run(eng, ch, engine.FP32)
            n_rr, ms_rr = eng.last_rerun()
            rel = np.abs(r32["score"] - s64) / np.maximum(s64, 1e-300)
            bad = rel > 1e-4
            badr = ~np.isclose(r32["rmsd"], r64["rmsd"], rtol=1e-4, atol=1e-6) | ~np.isclose(r32["tm"], r64["tm"], rtol=1e-4, atol=1e-9)
            marked = (r32["status"] & (engine.ST_TIE | engine.ST_FP64)) != 0
            diff_nc = r32["ncommon"] != r64["ncommon"]
            out = dict(config=name, pairs=len(s64), c=c, eps=eps, rerun=rerun, ms_fp32=round(ms32, 3), ms_fp64=round(ms64, 1),
                       marked=int(marked.sum()), marked_pct=round(100 * marked.mean(), 4), rerun_pairs=n_rr, rerun_ms=round(ms_rr, 3),
                       score_outside_1e4=int(bad.sum()), rmsd_tm_outside=int(badr.sum()), ncommon_differs=int(diff_nc.sum()),
                       outside_and_unmarked=int((bad & ~marked).sum()), ncommon_differs_unmarked=int((diff_nc & ~marked).sum()),
                       max_rel=float(rel.max()))
            um = np.nonzero((bad | diff_nc) & ~marked)[0]
            out['unmarked_pairs'] = [(int(pi[q]), int(pj[q]), float(rel[q])) for q in um[:12]]
            print(json.dumps(out), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
