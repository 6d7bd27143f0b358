"""flexible=True on the device: Protein.score_function(flexible=True) (tensor-only score matrix, multiple_alignment.py:323-326)
in the all-vs-all matrix and in the progressive alignment, Protein.mean_function(flexible=True) (coordinate-less nodes,
:359-360).  Golden vectors from the unmodified reference (oracle/gen_golden_flexible.py); the oracle for everything else."""
import os

import numpy as np
import pytest

from caretta_b200 import engine, synth
from caretta_b200 import multiple_alignment as MA
from oracle import oracle as O

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
FLEX = dict(flexible=True, gamma_tensor=7.0, gamma_coords=0.03)


def _case(g, name):
    L = g[f"{name}_lengths"]
    return synth.make_chains(len(L), list(L), 10, seed=int(g[f"{name}_seed"]), family_size=int(g[f"{name}_family"]))


def _proteins(ch, coordinates=True):
    return [MA.Protein(f"s{p}", ch.chain(p)[0], ch.chain(p)[1] if coordinates else None, "A" * ch.length(p)) for p in range(ch.n)]


@pytest.mark.parametrize("name", ["fam8", "ragged12", "short5"])
def test_flexible_pairwise_matrix_golden(name):
    g = np.load(os.path.join(G, "flexible.npz"))
    ch = _case(g, name)
    want = g[f"{name}_score"]
    # fp64 parity mode: the reference's expression order; CUDA exp is not glibc's to the last ulp (bound like the other fp64 tests)
    S64 = MA.MultipleAlignment(_proteins(ch), precision=engine.FP64).make_pairwise_matrix(dict(FLEX))
    np.testing.assert_allclose(S64, want, rtol=1e-11, atol=0)
    assert np.array_equal(S64 == 0, want == 0)
    # fp32 production mode, and proteins without coordinates (the reference never reads them with flexible=True)
    S32 = MA.MultipleAlignment(_proteins(ch, coordinates=False), precision=engine.FP32).make_pairwise_matrix(dict(FLEX))
    np.testing.assert_allclose(S32, want, rtol=1e-4, atol=1e-30)
    assert np.array_equal(S32, S32.T) and np.all(np.diag(S32) == 0)


def test_flexible_pair_list_by_products_are_zero():
    ch = synth.make_chains(6, [30, 41, 52, 37, 44, 60], 10, seed=77, family_size=3)
    eng = MA.get_engine()
    eng.set_chains(ch.coords, ch.tensors, ch.offsets)
    pi, pj = np.triu_indices(ch.n, 1)
    r = eng.pairwise_list(eng.params(7.0, 0.03, engine.FP64, flexible=True), pi, pj)
    want = O.pairwise_all_flexible(ch.tensors, ch.offsets, 7.0)
    np.testing.assert_allclose(r["score"], want[pi, pj], rtol=1e-11, atol=0)
    assert not r["rmsd"].any() and not r["tm"].any() and not r["ncommon"].any() and not r["status"].any()
    with pytest.raises(engine.CrtError):
        eng.pairwise_list(eng.params(7.0, 0.03, engine.FP64, flexible=True), pi, pj, want_paths=True)


@pytest.mark.parametrize("name", ["fam8", "ragged12", "short5"])
@pytest.mark.parametrize("tag", ["tt", "tf"])
@pytest.mark.parametrize("mode", ["pool", "level", "node"])
def test_flexible_multiple_align_golden(name, tag, mode, monkeypatch):
    g = np.load(os.path.join(G, "flexible.npz"))
    ch = _case(g, name)
    mean_flex = tag == "tt"
    monkeypatch.setenv("CARETTA_B200_NODE_BATCH", "0" if mode == "node" else "1")
    monkeypatch.setenv("CARETTA_B200_MSA_POOL", "1" if mode == "pool" else "0")
    S = g[f"{name}_score"]
    msa = MA.MultipleAlignment(_proteins(ch, coordinates=not mean_flex))
    aln = msa.multiple_align(np.max(S) - S, 1.0, 0.01, 1.0, 0.03, dict(FLEX), dict(flexible=mean_flex))
    assert np.array_equal(msa.tree, g[f"{name}_tree"])
    A = np.array([aln[f"s{p}"] for p in range(ch.n)])
    assert A.shape == g[f"{name}_{tag}_aln"].shape and np.array_equal(A, g[f"{name}_{tag}_aln"])
    fin = msa.final_sequences[-1]
    np.testing.assert_allclose(fin.tensors, g[f"{name}_{tag}_final_tensors"], rtol=0, atol=1e-12)
    assert np.array_equal(msa.final_consensus_weights[-1], g[f"{name}_{tag}_final_weights"])
    if mean_flex:
        assert fin.coordinates is None and all(s.coordinates is None for s in msa.final_sequences[ch.n:])
    else:
        np.testing.assert_allclose(fin.coordinates, g[f"{name}_{tag}_final_coords"], rtol=0, atol=1e-9)
    assert not msa.last_status.any()


def test_flexible_two_structures_and_bad_combination():
    g = np.load(os.path.join(G, "flexible.npz"))
    ch = _case(g, "two")
    msa = MA.MultipleAlignment(_proteins(ch, coordinates=False))
    aln = msa.multiple_align(None, 1.0, 0.01, 1.0, 0.03, dict(FLEX), dict(flexible=True))
    assert np.array_equal(np.array([aln["s0"], aln["s1"]]), g["two_tt_aln"])
    # coordinate-less nodes cannot be scored with flexible=False (the reference fails inside numba on the second level)
    ch3 = synth.make_chains(3, [20, 22, 25], 10, seed=5, family_size=3)
    with pytest.raises(ValueError):
        MA.MultipleAlignment(_proteins(ch3)).progressive_align(np.array([[0, 3], [1, 3], [3, 2]]), 1.0, 0.01, 1.0, 0.03,
                                                              dict(gamma_tensor=7.0), dict(flexible=True))


def test_flexible_node_vs_oracle():
    """One flexible node with non-trivial consensus weights against the oracle (both mean-function variants)."""
    ch = synth.make_chains(2, [73, 91], 10, seed=56, family_size=2)
    (t1, c1), (t2, c2) = ch.chain(0), ch.chain(1)
    rng = np.random.default_rng(3)
    w1, w2 = rng.integers(1, 5, (73, 1)).astype(np.float64), rng.integers(1, 4, (91, 1)).astype(np.float64)
    m1, m2 = 3 / (2 * (4 + 3)), 4 / (2 * (4 + 3))
    eng = MA.get_engine()
    for gc, mean_flex in ((engine.GC_FLEXIBLE, True), (engine.GC_FLEXIBLE_SCORE, False)):
        a1, a2, tm, cm, wm, sc, st = eng.progressive_node(t1, c1, w1, t2, c2, w2, m1, m2, 7.0, gc, 0.03, 1.0, 0.01)
        o1, o2, otm, ocm, owm, osc, _ = O.progressive_node(t1, c1, w1, t2, c2, w2, m1, m2, 7.0, 0.03, 0.03, 1.0, 0.01,
                                                           flexible_score=True, flexible_mean=mean_flex)
        assert np.array_equal(a1, o1) and np.array_equal(a2, o2) and st == 0
        np.testing.assert_allclose(sc, osc, rtol=1e-11)
        assert np.array_equal(tm, otm) and np.array_equal(wm, owm)
        if not mean_flex:
            np.testing.assert_allclose(cm, ocm, rtol=0, atol=1e-9)
