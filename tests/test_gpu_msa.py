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
    """The three ways to run the progressive alignment give bit-identical nodes: the device-resident pool (crt_msa_*, default),
    one crt_progressive_level call per tree level with host arrays, one crt_progressive_node call per node."""
    ch = synth.make_chains(48, list(np.random.default_rng(8).integers(40, 140, 48)), 10, seed=21, family_size=6)
    monkeypatch.setenv("CARETTA_B200_PRECISION", "fp64")
    msa = MA.StructureMultiple.from_chains(ch)
    S = msa.make_pairwise_matrix(dict(PARAMS))
    out = {}
    for mode, (batch, pool) in dict(pool=("1", "1"), level=("1", "0"), node=("0", "0")).items():
        monkeypatch.setenv("CARETTA_B200_NODE_BATCH", batch)
        monkeypatch.setenv("CARETTA_B200_MSA_POOL", pool)
        m = MA.StructureMultiple.from_chains(ch)
        aln = m.multiple_align(np.max(S) - S, 1.0, 0.01, 1.0, 0.03, dict(PARAMS), None)
        # read the nodes now: the pool belongs to the latest alignment on the engine
        out[mode] = (aln, m, [(s.name, s.tensors.copy(), s.coordinates.copy()) for s in m.final_sequences],
                     [np.array(w) for w in m.final_consensus_weights])
    a0, m0, seq0, w0 = out["node"]
    for mode in ("pool", "level"):
        a1, m1, seq1, w1 = out[mode]
        assert list(a1) == list(a0) and all(np.array_equal(a1[k], a0[k]) for k in a1), mode
        assert np.array_equal(m1.tree, m0.tree)
        assert len(seq1) == len(seq0) == 2 * ch.n - 1
        for (n1, t1, c1), (n0, t0, c0), x1, x0 in zip(seq1, seq0, w1, w0):
            assert n1 == n0 and np.array_equal(t1, t0) and np.array_equal(c1, c0) and np.array_equal(x1, x0), (mode, n1)
        assert np.array_equal(m1.last_status, m0.last_status)
        assert list(m1.final_alignments) == list(m0.final_alignments)
        for k in m0.final_alignments:
            assert list(m1.final_alignments[k]) == list(m0.final_alignments[k])
            assert all(np.array_equal(m1.final_alignments[k][x], m0.final_alignments[k][x]) for x in m0.final_alignments[k])
    # and the alignment is the oracle's (the CPU restatement of the reference's progressive_align) on the same tree
    seqs = [(f"s{p}", *ch.chain(p)) for p in range(ch.n)]
    want, _, _ = O.progressive_align(seqs, m0.tree, 1.0, 0.01, 1.0, 0.03, 7.0, 0.03)
    assert all(np.array_equal(a0[k], want[k]) for k in want)
    # a pool that is replaced by the next alignment hands its unread nodes over first: the reference keeps
    # final_sequences for as long as the MultipleAlignment lives, so reading m2 after m3 ran must work and give m2's nodes
    stale = out["pool"][1]
    fresh = MA.StructureMultiple.from_chains(ch)
    monkeypatch.setenv("CARETTA_B200_NODE_BATCH", "1")
    monkeypatch.setenv("CARETTA_B200_MSA_POOL", "1")
    m2 = MA.StructureMultiple.from_chains(ch)
    m2.multiple_align(np.max(S) - S, 1.0, 0.01, 1.0, 0.03, dict(PARAMS), None)
    m3 = MA.StructureMultiple.from_chains(ch)
    m3.multiple_align(np.max(S) - S, 1.0, 0.01, 1.0, 0.03, dict(PARAMS), None)
    last2, last3 = m2.final_sequences[-1], m3.final_sequences[-1]
    assert last2.name == "int-final" and np.array_equal(last2.tensors, last3.tensors) and np.array_equal(last2.tensors, seq0[-1][1])
    assert np.array_equal(np.array(m2.final_consensus_weights[-1]), w0[-1])
    assert stale.final_sequences[0].name == "s0" and fresh is not None


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


@pytest.mark.parametrize("d", [3, 13, 16])
def test_progressive_alignment_other_tensor_widths(d, monkeypatch):
    """Tensor widths other than 10 through the whole progressive alignment (pool mode) against the oracle's restatement."""
    ch = synth.make_chains(9, [44, 61, 38, 72, 55, 49, 66, 41, 58], d, seed=400 + d, family_size=3)
    monkeypatch.setenv("CARETTA_B200_PRECISION", "fp64")
    msa = MA.StructureMultiple.from_chains(ch)
    S = msa.make_pairwise_matrix(dict(PARAMS))
    aln = msa.multiple_align(np.max(S) - S, 1.0, 0.01, 1.0, 0.03, dict(PARAMS), None)
    seqs = [(f"s{p}", *ch.chain(p)) for p in range(ch.n)]
    want, fs, fw = O.progressive_align(seqs, msa.tree, 1.0, 0.01, 1.0, 0.03, 7.0, 0.03)
    assert list(aln) == list(want) and all(np.array_equal(aln[k], want[k]) for k in want)
    np.testing.assert_allclose(msa.final_sequences[-1].tensors, fs[-1][1], rtol=0, atol=1e-12)
    np.testing.assert_allclose(msa.final_sequences[-1].coordinates, fs[-1][2], rtol=0, atol=1e-9)
    assert np.array_equal(msa.final_consensus_weights[-1], fw[-1])


def test_level_is_cut_into_chunks_under_a_small_workspace(monkeypatch):
    """A level whose score matrices exceed the workspace budget runs in several chunks of nodes: same results."""
    rng = np.random.default_rng(6)
    eng = MA.get_engine()
    children, mults = [], []
    for q in range(11):
        n, m = int(rng.integers(20, 90)), int(rng.integers(20, 90))
        ch = synth.make_chains(2, [n, m], 10, seed=500 + q, family_size=2)
        (t1, c1), (t2, c2) = ch.chain(0), ch.chain(1)
        children.append(((t1, c1, np.full((n, 1), 1.0 + q % 3)), (t2, c2, np.full((m, 1), 2.0))))
        mults.append((0.2 + 0.01 * q, 0.3))
    monkeypatch.delenv("CARETTA_B200_LEVEL_CELLS", raising=False)
    want = eng.progressive_level(children, mults, 7.0, 0.03, 0.03, 1.0, 0.01)
    monkeypatch.setenv("CARETTA_B200_LEVEL_CELLS", "9000")                 # 2-3 nodes per chunk
    got = eng.progressive_level(children, mults, 7.0, 0.03, 0.03, 1.0, 0.01)
    for g_, w_ in zip(got, want):
        for x, y in zip(g_, w_):
            assert np.array_equal(np.asarray(x), np.asarray(y))
