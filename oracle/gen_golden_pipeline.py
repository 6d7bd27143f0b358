"""TEST INFRASTRUCTURE.  Golden FILES of a whole run of the UNMODIFIED reference after feature extraction
(align_from_structure_files, multiple_alignment.py:488-591, with full=True, write_fasta=True, write_matrix=True), on synthetic
proteins:  python oracle/gen_golden_pipeline.py -> tests/golden/pipeline.npz  (the bytes of result.fasta and of the four matrix
text files, plus the alignment)."""
import os
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_harness  # noqa: E402
from caretta_b200 import synth  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
CASE = dict(n=14, lengths=[70, 82, 64, 91, 77, 60, 85, 73, 66, 95, 58, 80, 71, 88], seed=301, family_size=7)
LETTERS = np.array(list("ACDEFGHIKLMNPQRSTVWY"))


def main():
    ma, dtw, sf, sup, helper, nj = ref_harness.load()
    ch = synth.make_chains(CASE["n"], CASE["lengths"], 10, seed=CASE["seed"], family_size=CASE["family_size"])
    rng = np.random.default_rng(CASE["seed"])
    seqs = ["".join(LETTERS[rng.integers(0, 20, int(L))]) for L in CASE["lengths"]]
    proteins = [ma.Protein(f"prot{p:02d}.pdb", ch.chain(p)[0].copy(), ch.chain(p)[1].copy(), seqs[p]) for p in range(ch.n)]
    # multiple_alignment.py:488-591 with the I/O-free steps only
    msa_class = ma.MultipleAlignment(proteins)
    sfp = dict(flexible=False, gamma_tensor=7., gamma_coords=0.03, verbose=False)
    D = msa_class.make_pairwise_matrix(score_function_params=sfp)
    D = D.max() - D
    out = {"lengths": np.array(CASE["lengths"]), "seed": CASE["seed"], "family": CASE["family_size"], "seqs": np.array(seqs)}
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        names = [s.name for s in msa_class.sequences]
        helper.write_distance_matrix(names, D, td / "distance_matrix_guide_tree.txt")
        alignment = msa_class.multiple_align(D, gap_open_penalty=1.0, gap_extend_penalty=0.01, consensus_weight=float(True), gamma_weight=1.,
                                             score_function_params=sfp, mean_function_params=dict(flexible=False, verbose=False))
        msa_class.write_alignment(td / "result.fasta")
        rmsd, coverage, tm = ma.make_rmsd_coverage_tm_matrix(alignment, msa_class.sequences, superpose_first=False)
        helper.write_distance_matrix(names, rmsd, td / "rmsd.txt")
        helper.write_distance_matrix(names, coverage, td / "coverage.txt")
        helper.write_distance_matrix(names, tm, td / "tm.txt")
        for f in ("distance_matrix_guide_tree.txt", "result.fasta", "rmsd.txt", "coverage.txt", "tm.txt"):
            out["file_" + f] = np.frombuffer((td / f).read_bytes(), dtype=np.uint8)
    out["aln"] = np.array([alignment[n] for n in names])
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(GOLD, "pipeline.npz"), **out)
    print("[gen-pipeline] alignment", out["aln"].shape, {k: len(v) for k, v in out.items() if k.startswith("file_")})


if __name__ == "__main__":
    main()
