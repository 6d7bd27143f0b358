"""TEST INFRASTRUCTURE.  Golden vectors for the progressive alignment, produced by running the
UNMODIFIED reference in the build container:  python oracle/gen_golden_msa.py  -> tests/golden/msa.npz.
For each case: all-vs-all score matrix -> max(S) - S -> neighbor_joining -> MultipleAlignment.multiple_align with the
reference defaults (gap open 1.0, extend 0.01, consensus_weight 1.0, gamma_weight 0.03, multiple_alignment.py:399-409,
490-492); stored: the guide tree and the final alignment matrix [N, A] (int64, -1 = gap)."""
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
    "fam8": dict(n=8, lengths=60, seed=101, family_size=8),
    "ragged12": dict(n=12, lengths=[40, 55, 70, 61, 48, 90, 33, 120, 77, 64, 52, 85], seed=102, family_size=4),
    "two": dict(n=2, lengths=[50, 58], seed=103, family_size=2),
    "mixed40": dict(n=40, lengths=[100 + (7 * k) % 41 for k in range(40)], seed=104, family_size=10),
}
PARAMS = dict(flexible=False, gamma_tensor=7.0, gamma_coords=0.03, verbose=False)


def main():
    ma, dtw, sf, sup, helper, nj = ref_harness.load()
    out = {"names": np.array(list(CASES))}
    for name, kw in CASES.items():
        t0 = time.time()
        ch = synth.make_chains(kw["n"], kw["lengths"], 10, seed=kw["seed"], family_size=kw["family_size"])
        P = ref_harness.proteins_from_chains(ma, ch)
        msa = ma.MultipleAlignment(P)
        S = msa.make_pairwise_matrix(dict(PARAMS))
        D = np.max(S) - S
        aln = msa.multiple_align(D, gap_open_penalty=1.0, gap_extend_penalty=0.01, consensus_weight=1.0, gamma_weight=0.03,
                                 score_function_params=dict(PARAMS), mean_function_params=dict(verbose=False))
        A = np.array([np.asarray(aln[p.name], dtype=np.int64) for p in P])
        out[f"{name}_aln"] = A
        out[f"{name}_lengths"] = ch.lengths
        out[f"{name}_seed"] = kw["seed"]
        out[f"{name}_family"] = kw["family_size"]
        out[f"{name}_score"] = S
        if msa.tree is not None:
            out[f"{name}_tree"] = msa.tree
            out[f"{name}_bl"] = msa.branch_lengths
            fin = msa.final_sequences[-1]
            out[f"{name}_final_tensors"] = fin.tensors
            out[f"{name}_final_coords"] = fin.coordinates
            out[f"{name}_final_weights"] = msa.final_consensus_weights[-1]
            # the bookkeeping dictionaries (final_alignments: node name -> {member name -> indices}), flattened in dict order
            keys, mem, lens, flat = [], [], [], []
            for k, dct in msa.final_alignments.items():
                for mname, arr in dct.items():
                    keys.append(k); mem.append(mname); lens.append(len(arr)); flat.append(np.asarray(arr, dtype=np.int64))
            out[f"{name}_fa_keys"], out[f"{name}_fa_members"] = np.array(keys), np.array(mem)
            out[f"{name}_fa_lens"], out[f"{name}_fa_flat"] = np.array(lens), np.concatenate(flat)
            out[f"{name}_fs_names"] = np.array([s.name for s in msa.final_sequences])
        print(f"[gen-msa] {name}: N={ch.n} alignment {A.shape}  {time.time() - t0:.1f}s")
    np.savez_compressed(os.path.join(GOLD, "msa.npz"), **out)


if __name__ == "__main__":
    main()
