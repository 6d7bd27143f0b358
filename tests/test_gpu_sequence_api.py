"""The SequenceBase / Protein interface of the reference (multiple_alignment.py:109-127, :312-389) on the device:
Protein.score_function / mean_function / get_mean_weights against golden vectors from the unmodified reference
(oracle/gen_golden_flexible.py), and the driver's generic path for sequences that are not Proteins (their own score_function /
mean_function on the host, the dynamic programming on the device)."""
import os
from dataclasses import dataclass

import numpy as np
import pytest

from caretta_b200 import engine, synth
from caretta_b200 import multiple_alignment as MA
from oracle import oracle as O

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", ["fn_a", "fn_short", "fn_b"])
@pytest.mark.parametrize("tag", ["rigid", "flex"])
def test_protein_methods_golden(name, tag, capsys):
    g = np.load(os.path.join(G, "flexible.npz"))
    lens, seed = [int(x) for x in g[f"{name}_lengths"]], int(g[f"{name}_seed"])
    flex = tag == "flex"
    ch = synth.make_chains(2, lens, 10, seed=seed, family_size=2)
    p1, p2 = (MA.Protein(f"s{k}", ch.chain(k)[0], None if flex else ch.chain(k)[1], "A" * lens[k]) for k in range(2))
    S = p1.score_function(p2, flexible=flex, gamma_tensor=7.0, gamma_coords=0.03, verbose=True)
    want = g[f"{name}_{tag}_S"]
    assert S.shape == want.shape and S.dtype == np.float64
    np.testing.assert_allclose(S, want, rtol=1e-8 if not flex else 1e-11, atol=1e-300)
    said = capsys.readouterr().out
    # <= 3 matched residues: the reference's message, no superposition (:337-342)
    assert ("Too few aligning positions for s0 and s1" in said) == (name == "fn_short" and not flex)
    a1, a2 = g[f"{name}_{tag}_aln"]
    # the affine DTW on the device, on OUR matrix, gives the reference's alignment
    d1, d2, _ = MA.dtw_align(np.arange(lens[0]), np.arange(lens[1]), S, gap_open_penalty=1.0, gap_extend_penalty=0.01)
    assert np.array_equal(d1, a1) and np.array_equal(d2, a2)
    node = p1.mean_function(p2, a1, a2, "int-x", flexible=flex, verbose=False)
    assert node.name == "int-x" and isinstance(node, MA.Protein) and len(node) == len(a1)
    assert np.array_equal(node.tensors, g[f"{name}_{tag}_tensors"])
    if flex:
        assert node.coordinates is None
    else:
        np.testing.assert_allclose(node.coordinates, g[f"{name}_{tag}_coords"], rtol=0, atol=1e-9)
    rng = np.random.default_rng(seed)
    w1, w2 = rng.integers(1, 5, (lens[0], 1)).astype(np.float64), rng.integers(1, 4, (lens[1], 1)).astype(np.float64)
    W = MA.get_mean_weights(w1, w2, a1, a2)
    assert W.shape == (len(a1), 1) and np.array_equal(W, g[f"{name}_{tag}_weights"])


def test_mean_function_argument_errors():
    eng = MA.get_engine()
    t = np.zeros((4, 10))
    with pytest.raises(engine.CrtError):
        eng.mean_function(t, None, t, None, [0, -1], [0, -1], flexible=True)            # a column with two gaps
    with pytest.raises(engine.CrtError):
        eng.mean_function(t, None, t, None, [0, 4], [0, 1], flexible=True)              # index out of range
    with pytest.raises(ValueError):
        eng.mean_function(t, None, t, None, [0, 1], [0], flexible=True)
    with pytest.raises(engine.CrtError):
        eng.mean_weights(np.ones(4), np.ones(4), [5], [0])
    assert eng.mean_weights(np.ones(4), np.ones(4), [], []).shape == (0, 1)


@dataclass
class FeatureSequence(MA.SequenceBase):
    """A SequenceBase that is not a Protein: no .tensors, its own score and mean functions (host code)."""
    name: str
    feat: np.ndarray

    def score_function(self, other, gamma=1.0):
        return O.rbf_matrix(self.feat, other.feat, gamma)

    def mean_function(self, other, aln_1, aln_2, name_int):
        out = np.zeros((len(aln_1), self.feat.shape[1]))
        for i, (x, y) in enumerate(zip(aln_1, aln_2)):
            out[i] = other.feat[y] if x == -1 else (self.feat[x] if y == -1 else (self.feat[x] + other.feat[y]) / 2)
        return FeatureSequence(name_int, out)

    def __len__(self):
        return self.feat.shape[0]

    def __str__(self):
        return "X" * len(self)


@pytest.mark.parametrize("name", ["fam8", "ragged12", "short5"])
def test_generic_sequence_base_driver(name):
    """The tensor Gaussian as a user-defined score function is the reference's flexible=True run: same matrix (bit for bit: the
    Smith-Waterman score on the device is exact float64 add / max), same tree, same alignment, same consensus."""
    g = np.load(os.path.join(G, "flexible.npz"))
    L = g[f"{name}_lengths"]
    ch = synth.make_chains(len(L), list(L), 10, seed=int(g[f"{name}_seed"]), family_size=int(g[f"{name}_family"]))
    seqs = [FeatureSequence(f"s{p}", ch.chain(p)[0]) for p in range(ch.n)]
    msa = MA.MultipleAlignment(seqs)
    S = msa.make_pairwise_matrix(dict(gamma=7.0))
    assert np.array_equal(S, g[f"{name}_score"])
    aln = msa.multiple_align(np.max(S) - S, 1.0, 0.01, 1.0, 0.03, dict(gamma=7.0), None)
    assert np.array_equal(msa.tree, g[f"{name}_tree"])
    assert sorted(aln) == sorted(f"s{p}" for p in range(ch.n))          # dict order = tree order, like the reference's
    assert np.array_equal(np.array([aln[f"s{p}"] for p in range(ch.n)]), g[f"{name}_tt_aln"])
    assert np.array_equal(msa.final_sequences[-1].feat, g[f"{name}_tt_final_tensors"])
    assert np.array_equal(msa.final_consensus_weights[-1], g[f"{name}_tt_final_weights"])
    assert list(msa.final_alignments)[-1] == "int-final" and len(msa.final_sequences) == 2 * ch.n - 1


def test_generic_two_sequences_and_batches():
    g = np.load(os.path.join(G, "flexible.npz"))
    L = g["two_lengths"]
    ch = synth.make_chains(2, list(L), 10, seed=int(g["two_seed"]), family_size=int(g["two_family"]))
    seqs = [FeatureSequence(f"s{p}", ch.chain(p)[0]) for p in range(2)]
    aln = MA.MultipleAlignment(seqs).multiple_align(None, 1.0, 0.01, 1.0, 0.03, dict(gamma=7.0), None)
    assert np.array_equal(np.array([aln["s0"], aln["s1"]]), g["two_tt_aln"])
    # several device batches of matrices give the same matrix as one
    ch = synth.make_chains(7, [30, 41, 52, 37, 44, 60, 25], 10, seed=9, family_size=7)
    seqs = [FeatureSequence(f"s{p}", ch.chain(p)[0]) for p in range(ch.n)]
    one = MA.MultipleAlignment(seqs)._pairwise_matrix_generic(dict(gamma=7.0))
    many = MA.MultipleAlignment(seqs)._pairwise_matrix_generic(dict(gamma=7.0), batch_bytes=40000)
    assert np.array_equal(one, many) and np.array_equal(one, O.pairwise_all_flexible(ch.tensors, ch.offsets, 7.0))


def test_make_score_matrix_is_the_reference_gaussian():
    """score_functions.make_score_matrix with get_gaussian_score on the device against the oracle's RBF (pinned bit-exact on the
    reference's): any feature width, both gammas of the path; agreement to the last ulp of CUDA's exp (<= 1e-14 relative)."""
    rng = np.random.default_rng(12)
    for (n, m, k, gamma) in ((37, 51, 10, 7.0), (64, 20, 3, 0.03), (5, 9, 1, 1.0), (1, 1, 16, 0.5)):
        x, y = rng.normal(0, 0.4, (n, k)), rng.normal(0, 0.4, (m, k))
        S = MA.make_score_matrix(x, y, MA.get_gaussian_score, gamma)
        assert S.shape == (n, m)
        np.testing.assert_allclose(S, O.rbf_matrix(x, y, gamma), rtol=1e-14, atol=0)
    assert MA.get_gaussian_score(np.array([1.0, 2.0, 3.0]), np.array([1.5, 2.0, 2.0]), 0.03) == pytest.approx(np.exp(-0.03 * 1.25), rel=1e-15)
    with pytest.raises(NotImplementedError):
        MA.make_score_matrix(x, y, lambda a, b, g: 0.0, 1.0)
