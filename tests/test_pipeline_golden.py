"""The files of a whole reference run after feature extraction (tests/golden/pipeline.npz, written by the unmodified reference:
oracle/gen_golden_pipeline.py): first the oracle's composition of the path (CPU), then the product through the C ABI (GPU)."""
import os

import numpy as np
import pytest

from caretta_b200 import synth

G = os.path.join(os.path.dirname(__file__), "golden")
FILES = ("distance_matrix_guide_tree.txt", "result.fasta", "rmsd.txt", "coverage.txt", "tm.txt")


def _inputs():
    g = np.load(os.path.join(G, "pipeline.npz"))
    ch = synth.make_chains(len(g["lengths"]), list(g["lengths"]), 10, seed=int(g["seed"]), family_size=int(g["family"]))
    return g, ch, [str(x) for x in g["names"]], [str(x) for x in g["seqs"]]


def test_oracle_reproduces_the_reference_files():
    from oracle import oracle as O
    g, ch, names, seqs = _inputs()
    S = O.pairwise_all(ch.coords, ch.tensors, ch.offsets)
    D = S.max() - S
    assert O.format_matrix(names, D) == g["file_distance_matrix_guide_tree.txt"].tobytes()
    tree, _ = O.neighbor_joining(D)
    aln, _, _ = O.progressive_align([(names[p],) + ch.chain(p) for p in range(ch.n)], tree, 1.0, 0.01, 1.0, 1.0, 7.0, 0.03)
    A = np.array([aln[n] for n in names])
    assert np.array_equal(A, g["aln"])
    assert O.format_fasta(names, seqs, A) == g["file_result.fasta"].tobytes()
    r, c, t, bad = O.rmsd_cov_tm(A, ch.coords, ch.offsets)
    assert bad == 0
    for f, M in (("rmsd.txt", r), ("coverage.txt", c), ("tm.txt", t)):
        assert O.format_matrix(names, M) == g["file_" + f].tobytes(), f


@pytest.mark.gpu
@pytest.mark.parametrize("batch", ["1", "0"])
def test_align_from_proteins_writes_the_reference_files(tmp_path, monkeypatch, batch):
    from caretta_b200 import multiple_alignment as MA
    g, ch, names, seqs = _inputs()
    monkeypatch.setenv("CARETTA_B200_PRECISION", "fp64")
    monkeypatch.setenv("CARETTA_B200_NODE_BATCH", batch)
    proteins = [MA.Protein(names[p], ch.chain(p)[0].copy(), ch.chain(p)[1].copy(), seqs[p]) for p in range(ch.n)]
    msa, out = MA.align_from_proteins(proteins, output_folder=tmp_path / "caretta_results", write_fasta=True, write_matrix=True)
    assert np.array_equal(np.array([msa.alignment[n] for n in names]), g["aln"])
    got = {"result.fasta": out.fasta_file.read_bytes()}
    for f in FILES:
        if f != "result.fasta":
            got[f] = (out.matrix_folder / f).read_bytes()
    for f in FILES:
        assert got[f] == g["file_" + f].tobytes(), f


@pytest.mark.gpu
def test_align_from_proteins_fp32_same_alignment(tmp_path, monkeypatch):
    """Production precision: the guide matrix differs in the 7th digit, the alignment of this family set is the same."""
    from caretta_b200 import multiple_alignment as MA
    g, ch, names, seqs = _inputs()
    monkeypatch.setenv("CARETTA_B200_PRECISION", "fp32")
    proteins = [MA.Protein(names[p], ch.chain(p)[0].copy(), ch.chain(p)[1].copy(), seqs[p]) for p in range(ch.n)]
    msa, out = MA.align_from_proteins(proteins, output_folder=tmp_path / "r", write_fasta=True)
    assert np.array_equal(np.array([msa.alignment[n] for n in names]), g["aln"])
    assert out.fasta_file.read_bytes() == g["file_result.fasta"].tobytes()


@pytest.mark.gpu
def test_align_from_proteins_fast_mode_and_class_file(tmp_path, monkeypatch):
    """--fast (multiple_alignment.py:503-511): the guide matrix is the Bray-Curtis distance of shapemer counts (here: given index
    arrays), everything after it as in the full run, checked against the oracle's composition; --class (:557-559): the pickled
    object carries the plain dictionaries / lists of the reference."""
    import pickle
    from caretta_b200 import multiple_alignment as MA
    from oracle import oracle as O
    g, ch, names, seqs = _inputs()
    monkeypatch.setenv("CARETTA_B200_PRECISION", "fp64")
    rng = np.random.default_rng(31)
    K = 64
    # family members share most of their shapemers, so the guide tree is not degenerate
    base = [rng.integers(0, K, 40) for _ in range(1 + ch.n // int(g["family"]))]
    idx = [np.concatenate([base[p // int(g["family"])], rng.integers(0, K, 6)]) for p in range(ch.n)]
    proteins = [MA.Protein(names[p], ch.chain(p)[0].copy(), ch.chain(p)[1].copy(), seqs[p]) for p in range(ch.n)]
    msa, out = MA.align_from_proteins(proteins, output_folder=tmp_path / "fast", write_fasta=True, write_matrix=True, full=False,
                                      shapemer_indices=idx, alphabet_size=K, write_class=True)
    counts = O.count_matrix(idx, K)
    D = O.braycurtis(counts, counts)
    assert np.array_equal(msa.pairwise_distance_matrix, D)
    assert (out.matrix_folder / "distance_matrix_guide_tree.txt").read_bytes() == O.format_matrix(names, D)
    tree, bl = O.neighbor_joining(D)
    assert np.array_equal(msa.tree, tree) and np.array_equal(msa.branch_lengths, bl)
    want, _, _ = O.progressive_align([(names[p],) + ch.chain(p) for p in range(ch.n)], tree, 1.0, 0.01, 1.0, 1.0, 7.0, 0.03)
    A = np.array([want[n] for n in names])
    assert np.array_equal(np.array([msa.alignment[n] for n in names]), A)
    assert out.fasta_file.read_bytes() == O.format_fasta(names, seqs, A)
    clone = pickle.loads(out.class_file.read_bytes())
    assert type(clone.final_alignments) is dict and type(clone.final_sequences) is list
    assert [s.name for s in clone.final_sequences][-1] == "int-final" and len(clone.final_sequences) == 2 * ch.n - 1
    assert all(np.array_equal(clone.alignment[n], msa.alignment[n]) for n in names)
    assert np.array_equal(clone.tree, tree)
    with pytest.raises(ValueError):
        MA.align_from_proteins(proteins, full=False)
