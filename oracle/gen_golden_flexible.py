"""TEST INFRASTRUCTURE.  Golden vectors for flexible=True (tensor-only scoring, multiple_alignment.py:323-326, and
coordinate-less consensus nodes, :359-360), produced by running the UNMODIFIED reference in the build container:
    python oracle/gen_golden_flexible.py  -> tests/golden/flexible.npz
For each case: make_pairwise_matrix(flexible=True) -> max(S) - S -> multiple_align with score flexible=True and the mean
function flexible=True ("tt") or False ("tf"); stored: the score matrix, the guide tree, the final alignments and the final
consensus node."""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_harness  # noqa: E402
from caretta_b200 import synth  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
CASES = {
    "fam8": dict(n=8, lengths=60, seed=201, family_size=8),
    "ragged12": dict(n=12, lengths=[40, 55, 70, 61, 48, 90, 33, 120, 77, 64, 52, 85], seed=202, family_size=4),
    "two": dict(n=2, lengths=[50, 58], seed=203, family_size=2),
    "short5": dict(n=5, lengths=[1, 2, 3, 7, 12], seed=204, family_size=5),
}
PARAMS = dict(flexible=True, gamma_tensor=7.0, gamma_coords=0.03, verbose=False)


def main():
    ma, dtw, sf, sup, helper, nj = ref_harness.load()
    out = {"names": np.array(list(CASES))}
    for name, kw in CASES.items():
        t0 = time.time()
        ch = synth.make_chains(kw["n"], kw["lengths"], 10, seed=kw["seed"], family_size=kw["family_size"])
        out[f"{name}_lengths"], out[f"{name}_seed"], out[f"{name}_family"] = ch.lengths, kw["seed"], kw["family_size"]
        P = ref_harness.proteins_from_chains(ma, ch)
        S = ma.MultipleAlignment(P).make_pairwise_matrix(dict(PARAMS))
        out[f"{name}_score"] = S
        D = np.max(S) - S
        for tag, mean_flex in (("tt", True), ("tf", False)):
            msa = ma.MultipleAlignment(ref_harness.proteins_from_chains(ma, ch))
            aln = msa.multiple_align(D, gap_open_penalty=1.0, gap_extend_penalty=0.01, consensus_weight=1.0, gamma_weight=0.03,
                                     score_function_params=dict(PARAMS), mean_function_params=dict(flexible=mean_flex, verbose=False))
            out[f"{name}_{tag}_aln"] = np.array([np.asarray(aln[p.name], dtype=np.int64) for p in P])
            if msa.tree is not None:
                out[f"{name}_tree"], out[f"{name}_bl"] = msa.tree, msa.branch_lengths
                fin = msa.final_sequences[-1]
                out[f"{name}_{tag}_final_tensors"] = fin.tensors
                assert (fin.coordinates is None) == mean_flex
                if not mean_flex:
                    out[f"{name}_{tag}_final_coords"] = fin.coordinates
                out[f"{name}_{tag}_final_weights"] = msa.final_consensus_weights[-1]
        print(f"[gen-flexible] {name}: N={ch.n} alignment {out[f'{name}_tt_aln'].shape}  {time.time() - t0:.1f}s")
    # Protein.score_function / mean_function / get_mean_weights called directly (:321-383, :73-82), both flexible settings
    for name, (lens, seed) in dict(fn_a=([73, 91], 211), fn_short=([3, 2], 212), fn_b=([140, 37], 213)).items():
        ch = synth.make_chains(2, lens, 10, seed=seed, family_size=2)
        out[f"{name}_lengths"], out[f"{name}_seed"] = np.array(lens), seed
        p1, p2 = ref_harness.proteins_from_chains(ma, ch)
        rng = np.random.default_rng(seed)
        w1, w2 = rng.integers(1, 5, (lens[0], 1)).astype(np.float64), rng.integers(1, 4, (lens[1], 1)).astype(np.float64)
        for tag, flex in (("rigid", False), ("flex", True)):
            S = p1.score_function(p2, flexible=flex, gamma_tensor=7.0, gamma_coords=0.03, verbose=False)
            a1, a2, _ = dtw.dtw_align(np.arange(lens[0]), np.arange(lens[1]), S, gap_open_penalty=1.0, gap_extend_penalty=0.01)
            node = p1.mean_function(p2, a1, a2, "int-x", flexible=flex, verbose=False)
            out[f"{name}_{tag}_S"], out[f"{name}_{tag}_aln"] = S, np.array([a1, a2])
            out[f"{name}_{tag}_tensors"] = node.tensors
            if not flex:
                out[f"{name}_{tag}_coords"] = node.coordinates
            out[f"{name}_{tag}_weights"] = ma.get_mean_weights(w1, w2, a1, a2)
        print(f"[gen-flexible] {name}: {lens}")
    np.savez_compressed(os.path.join(GOLD, "flexible.npz"), **out)


if __name__ == "__main__":
    main()
