"""Progressive alignment on the device against the golden vectors produced by the unmodified
reference (oracle/gen_golden_msa.py): guide tree bit-identical, final alignment identical, final consensus node within
1e-9 (the reference's superposition goes through LAPACK/BLAS); and one node against the pinned oracle."""
import os

import numpy as np
import pytest

from caretta_b200 import engine, synth
from caretta_b200 import multiple_alignment as MA
from oracle import oracle as O

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
PARAMS = dict(gamma_tensor=7.0, gamma_coords=0.03)


def _case(g, name):
    L = g[f"{name}_lengths"]
    return synth.make_chains(len(L), list(L), 10, seed=int(g[f"{name}_seed"]), family_size=int(g[f"{name}_family"]))


@pytest.mark.parametrize("name", ["fam8", "ragged12", "mixed40"])
def test_multiple_align_golden(name):
    g = np.load(os.path.join(G, "msa.npz"))
    ch = _case(g, name)
    os.environ["CARETTA_B200_PRECISION"] = "fp64"
    try:
        msa = MA.StructureMultiple.from_chains(ch)
        S = msa.make_pairwise_matrix(dict(PARAMS))
        np.testing.assert_allclose(S, g[f"{name}_score"], rtol=1e-11)
        # the guide tree is checked on the reference's own matrix (bit-identical input -> bit-identical tree)
        Sref = g[f"{name}_score"]
        aln = msa.multiple_align(np.max(Sref) - Sref, 1.0, 0.01, 1.0, 0.03, dict(PARAMS), None)
    finally:
        os.environ.pop("CARETTA_B200_PRECISION")
    assert np.array_equal(msa.tree, g[f"{name}_tree"]) and np.array_equal(msa.branch_lengths, g[f"{name}_bl"])
    A = np.array([aln[f"s{p}"] for p in range(ch.n)])
    assert A.shape == g[f"{name}_aln"].shape and np.array_equal(A, g[f"{name}_aln"])
    fin = msa.final_sequences[-1]
    np.testing.assert_allclose(fin.tensors, g[f"{name}_final_tensors"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(fin.coordinates, g[f"{name}_final_coords"], rtol=0, atol=1e-9)
    assert np.array_equal(msa.final_consensus_weights[-1], g[f"{name}_final_weights"])


def test_two_structures_golden():
    g = np.load(os.path.join(G, "msa.npz"))
    ch = _case(g, "two")
    msa = MA.StructureMultiple.from_chains(ch)
    aln = msa.multiple_align(None, 1.0, 0.01, 1.0, 0.03, dict(PARAMS), None)
    assert np.array_equal(np.array([aln["s0"], aln["s1"]]), g["two_aln"])


def test_node_vs_oracle():
    """One node with non-trivial consensus weights and multipliers against the oracle's restatement."""
    ch = synth.make_chains(2, [73, 91], 10, seed=55, family_size=2)
    (t1, c1), (t2, c2) = ch.chain(0), ch.chain(1)
    rng = np.random.default_rng(3)
    w1, w2 = rng.integers(1, 5, (73, 1)).astype(np.float64), rng.integers(1, 4, (91, 1)).astype(np.float64)
    m1, m2 = 3 / (2 * (4 + 3)), 4 / (2 * (4 + 3))
    eng = MA.get_engine()
    a1, a2, tm, cm, wm, sc, st = eng.progressive_node(t1, c1, w1, t2, c2, w2, m1, m2, 7.0, 0.03, 0.03, 1.0, 0.01)
    S = O.score_matrix(t1, c1, t2, c2, 7.0, 0.03) + O.rbf_matrix(w1 * m1, w2 * m2, 0.03)
    o1, o2, osc = O.dtw_align(S, 1.0, 0.01)
    assert np.array_equal(a1, o1) and np.array_equal(a2, o2)
    np.testing.assert_allclose(sc, osc, rtol=1e-11)
    otm, ocm = O.mean_function(t1, c1, t2, c2, o1, o2)
    np.testing.assert_allclose(tm, otm, rtol=0, atol=1e-13)
    np.testing.assert_allclose(cm, ocm, rtol=0, atol=1e-9)
    assert np.array_equal(wm, O.mean_weights(w1, w2, o1, o2))


def test_level_batching_equals_node_by_node(monkeypatch):
    """crt_progressive_level (all nodes of a tree level in one call) gives bit-identical nodes to crt_progressive_node."""
    ch = synth.make_chains(48, list(np.random.default_rng(8).integers(40, 140, 48)), 10, seed=21, family_size=6)
    monkeypatch.setenv("CARETTA_B200_PRECISION", "fp64")
    msa = MA.StructureMultiple.from_chains(ch)
    S = msa.make_pairwise_matrix(dict(PARAMS))
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("CARETTA_B200_NODE_BATCH", mode)
        m = MA.StructureMultiple.from_chains(ch)
        aln = m.multiple_align(np.max(S) - S, 1.0, 0.01, 1.0, 0.03, dict(PARAMS), None)
        out[mode] = (aln, m)
    a1, m1 = out["1"]
    a0, m0 = out["0"]
    assert list(a1) == list(a0) and all(np.array_equal(a1[k], a0[k]) for k in a1)
    assert np.array_equal(m1.tree, m0.tree)
    for s1, s0, w1, w0 in zip(m1.final_sequences, m0.final_sequences, m1.final_consensus_weights, m0.final_consensus_weights):
        assert s1.name == s0.name and np.array_equal(s1.tensors, s0.tensors) and np.array_equal(s1.coordinates, s0.coordinates)
        assert np.array_equal(w1, w0)
    assert np.array_equal(m1.last_status, m0.last_status)
    # and the alignment is the oracle's (the CPU restatement of the reference's progressive_align) on the same tree
    seqs = [(f"s{p}", *ch.chain(p)) for p in range(ch.n)]
    want, _, _ = O.progressive_align(seqs, m1.tree, 1.0, 0.01, 1.0, 0.03, 7.0, 0.03)
    assert all(np.array_equal(a1[k], want[k]) for k in want)


def test_level_call_with_mixed_shapes():
    """One level call with very different node shapes (1-residue chains, a multi-strip DTW, <= 3 common positions)."""
    rng = np.random.default_rng(4)
    lens = [(1, 1), (1, 37), (3, 2), (150, 140), (64, 200), (33, 33)]
    eng = MA.get_engine()
    children, mults, want = [], [], []
    for q, (n, m) in enumerate(lens):
        ch = synth.make_chains(2, [n, m], 10, seed=300 + q, family_size=2)
        (t1, c1), (t2, c2) = ch.chain(0), ch.chain(1)
        w1, w2 = rng.integers(1, 4, (n, 1)).astype(np.float64), rng.integers(1, 4, (m, 1)).astype(np.float64)
        mu = (0.25 + 0.05 * q, 0.4 - 0.03 * q)
        children.append(((t1, c1, w1), (t2, c2, w2)))
        mults.append(mu)
        want.append(eng.progressive_node(t1, c1, w1, t2, c2, w2, mu[0], mu[1], 7.0, 0.03, 0.03, 1.0, 0.01))
    got = eng.progressive_level(children, mults, 7.0, 0.03, 0.03, 1.0, 0.01)
    for g_, w_ in zip(got, want):
        for x, y in zip(g_, w_):
            assert np.array_equal(np.asarray(x), np.asarray(y))
    o = O.progressive_node(*children[3][0], *children[3][1], *mults[3], 7.0, 0.03, 0.03, 1.0, 0.01)
    assert np.array_equal(got[3][0], o[0]) and np.array_equal(got[3][1], o[1])
    np.testing.assert_allclose(got[3][3], o[3], rtol=0, atol=1e-9)
