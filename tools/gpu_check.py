"""Diagnostic run on a GPU box: parity statistics against the golden fixtures + a first timing.  Prints, does not assert."""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from caretta_b200 import synth, engine

G = os.path.join(ROOT, "tests", "golden")


def cols(a1, a2):
    return set((int(x), int(y)) for x, y in zip(a1, a2) if x >= 0 and y >= 0)


def compare(tag, res, gold_a1, gold_a2, gold_off, gold, n):
    off = res["aln_off"]
    exact = 0; tot = same = 0; bad = []
    for q in range(n):
        a1 = res["aln1"][off[q]:off[q + 1]]; a2 = res["aln2"][off[q]:off[q + 1]]
        r1 = gold_a1[gold_off[q]:gold_off[q + 1]].astype(np.int64); r2 = gold_a2[gold_off[q]:gold_off[q + 1]].astype(np.int64)
        if len(a1) == len(r1) and np.array_equal(a1, r1) and np.array_equal(a2, r2):
            exact += 1
        elif len(bad) < 3:
            bad.append(q)
        cr, cg = cols(r1, r2), cols(a1, a2)
        tot += len(cr); same += len(cr & cg)
    rel = np.abs(res["score"] - gold["score"]) / np.maximum(np.abs(gold["score"]), 1e-300)
    print(f"[{tag}] paths exact {exact}/{n}  identical columns {same}/{tot} = {same / max(tot, 1):.6f}  "
          f"score rel err max {rel.max():.3e} median {np.median(rel):.3e}  "
          f"ncommon equal {(res['ncommon'] == gold['ncommon']).mean():.4f}  "
          f"rmsd max abs {np.abs(res['rmsd'] - gold['rmsd']).max():.3e}  tm max abs {np.abs(res['tm'] - gold['tm']).max():.3e}")
    return bad


def main():
    eng = engine.Engine()
    print("device", eng.device_info())
    small = np.load(os.path.join(G, "pairs_small.npz"))
    ch = synth.make_chains(len(small["lengths"]), small["lengths"], 10, seed=11, family_size=4)
    eng.set_chains(ch.coords, ch.tensors, ch.offsets)
    for prec, name in [(engine.FP64, "fp64"), (engine.FP32, "fp32")]:
        prm = eng.params(precision=prec)
        res = eng.pairwise_list(prm, small["pi"], small["pj"], want_paths=True)
        bad = compare(f"small {name}", res, small["aln1"], small["aln2"], small["aln_off"], small, len(small["pi"]))
        for q in bad:
            o = res["aln_off"]; go = small["aln_off"]
            print("   first mismatch pair", q, (small["pi"][q], small["pj"][q]), "status", res["status"][q])
            print("     got ", res["aln1"][o[q]:o[q + 1]][:20], res["aln2"][o[q]:o[q + 1]][:20])
            print("     want", small["aln1"][go[q]:go[q + 1]][:20], small["aln2"][go[q]:go[q + 1]][:20])
        S = eng.pairwise_all(prm)
        print(f"   pairwise_all vs golden matrix: max rel {np.max(np.abs(S - small['score_matrix']) / np.maximum(small['score_matrix'], 1e-300) * (small['score_matrix'] > 0)):.3e}  sym {np.array_equal(S, S.T)} diag0 {np.all(np.diag(S) == 0)}")
    # ---- C2
    g = np.load(os.path.join(G, "c2_full.npz"))
    ch = synth.config("C2")
    eng.set_chains(ch.coords, ch.tensors, ch.offsets)
    pi, pj = np.triu_indices(ch.n, 1)
    for prec, name in [(engine.FP64, "fp64"), (engine.FP32, "fp32")]:
        prm = eng.params(precision=prec)
        t = time.time()
        res = eng.pairwise_list(prm, pi, pj, want_paths=True)
        dt = time.time() - t
        compare(f"C2 {name}", res, g["aln1"], g["aln2"], g["aln_off"], g, len(pi))
        print(f"   wall {dt:.3f}s  device {eng.last_elapsed_ms():.2f} ms  launches {eng.last_launches()}")
    # ---- timing on C3
    peak, ms = eng.fp32_peak()
    print(f"fp32 FFMA peak {peak / 1e12:.2f} T lane-FFMA/s ({ms:.2f} ms)")
    ch = synth.config("C3")
    t = time.time(); eng.set_chains(ch.coords, ch.tensors, ch.offsets); print("set_chains C3", time.time() - t)
    for prec, name in [(engine.FP32, "fp32"), (engine.FP64, "fp64")]:
        prm = eng.params(precision=prec)
        for rep in range(2 if prec == engine.FP32 else 1):
            t = time.time(); eng.pairwise_shard(prm, 0, 1); dt = time.time() - t
            ms = eng.last_elapsed_ms(); cu = eng.last_cell_updates()
            npairs = ch.n * (ch.n - 1) // 2
            print(f"C3 {name} rep{rep}: wall {dt:.3f}s device {ms:.1f} ms  pairs/s {npairs / (ms * 1e-3):.0f}  GCUPS {cu / (ms * 1e-3) / 1e9:.1f}  "
                  f"lane-instr frac of peak {(cu / 2 * 36) / (ms * 1e-3) / peak:.3f}")
    r = eng.fetch(10)
    print("sample scores", r["score"][:5])


if __name__ == "__main__":
    main()
