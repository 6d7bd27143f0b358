"""DP kernels in isolation (affine DTW, Smith-Waterman with any gap) and the MSA RMSD/coverage/TM kernel against
the reference-generated goldens (bit-exact paths / scores) and the oracle on larger ragged problems."""
import os

import numpy as np
import pytest

from caretta_b200 import engine, synth
from caretta_b200 import multiple_alignment as MA
from oracle import oracle as O

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def eng():
    e = engine.Engine()
    yield e
    e.close()


def test_dtw_golden_bit_exact(eng):
    g = np.load(os.path.join(G, "dtw.npz"))
    mats, want = [], []
    for q in range(int(g["n_cases"])):
        for go, ge in [(1.0, 0.01), (0.5, 0.1), (0.0, 0.0)]:
            key = f"c{q}_{go}_{ge}"
            res = eng.dtw_align_batch([g[f"c{q}_S"]], go, ge)[0]
            assert res[0].tolist() == g[f"{key}_a1"].tolist(), key
            assert res[1].tolist() == g[f"{key}_a2"].tolist(), key
            assert res[2] == float(g[f"{key}_sc"]), key
    for q in range(6):
        mats.append(g[f"r{q}_S"]); want.append((g[f"r{q}_a1"], g[f"r{q}_a2"], float(g[f"r{q}_sc"])))
    for (a1, a2, sc), (w1, w2, wsc) in zip(eng.dtw_align_batch(mats, 1.0, 0.01), want):       # one batched call
        assert a1.tolist() == w1.tolist() and a2.tolist() == w2.tolist() and sc == wsc
    for (a1, a2, sc, st), q in zip(eng.sw_align_batch(mats, 0.0), range(6)):
        assert a1.tolist() == g[f"r{q}_sw_a1"].tolist() and a2.tolist() == g[f"r{q}_sw_a2"].tolist()
        assert sc == float(g[f"r{q}_sw_sc"]) and st == 0


def test_kat(eng):
    k = np.load(os.path.join(G, "kat.npz"))
    for q in range(7):
        S = k[f"dtw{q}_S"]
        go, ge = (1.0, 0.01) if q == 0 else (float(k[f"dtw{q}_go"]), float(k[f"dtw{q}_ge"]))
        a1, a2, sc = MA.dtw_align(np.arange(S.shape[0]), np.arange(S.shape[1]), S, go, ge)
        assert a1.dtype == np.int64
        assert a1.tolist() == k[f"dtw{q}_a1"].tolist() and a2.tolist() == k[f"dtw{q}_a2"].tolist() and sc == float(k[f"dtw{q}_sc"])
    for q in range(5):
        S = k[f"sw{q}_S"]
        a1, a2, sc = MA.smith_waterman(np.arange(S.shape[0]), np.arange(S.shape[1]), S, 0.0)
        assert a1.tolist() == k[f"sw{q}_a1"].tolist() and a2.tolist() == k[f"sw{q}_a2"].tolist() and sc == float(k[f"sw{q}_sc"])
        assert MA.smith_waterman_score(np.arange(S.shape[0]), np.arange(S.shape[1]), S) == float(k[f"sw{q}_score_only"])
    with pytest.raises(TypeError):                              # the reference raises on an all-zero matrix
        MA.smith_waterman(np.arange(3), np.arange(3), np.zeros((3, 3)), 0.0)
    assert MA.smith_waterman_score(np.arange(3), np.arange(3), np.zeros((3, 3))) == 0.0


def test_large_ragged_vs_oracle(eng):
    """Multi-strip shapes (m > 128), real Gaussian score matrices, several gap settings, one batch."""
    ch = synth.make_chains(6, [150, 333, 97, 401, 260, 129], 10, seed=77, family_size=3)
    mats = []
    for i, j in [(0, 1), (1, 3), (2, 5), (3, 4), (4, 0), (5, 1)]:
        ti, ci = ch.chain(i); tj, cj = ch.chain(j)
        mats.append(O.rbf_matrix(ti, tj, 7.0) + O.rbf_matrix(np.ones((len(ti), 1)), np.ones((len(tj), 1)), 1.0))
    for go, ge in [(1.0, 0.01), (0.3, 0.3)]:
        for (a1, a2, sc), S in zip(eng.dtw_align_batch(mats, go, ge), mats):
            w1, w2, wsc = O.dtw_align(S, go, ge)
            assert a1.tolist() == w1.tolist() and a2.tolist() == w2.tolist() and sc == wsc
    for gap in (0.0, 0.05, 0.4):
        for (a1, a2, sc, st), S in zip(eng.sw_align_batch(mats, gap), mats):
            w1, w2, wsc = O.smith_waterman(S, gap)
            assert a1.tolist() == w1.tolist() and a2.tolist() == w2.tolist() and sc == wsc
            assert eng.sw_align_batch([S], gap, want_paths=False)[0][2] == O.smith_waterman_score(S, gap)


def test_rmsd_cov_tm_golden(eng):
    g = np.load(os.path.join(G, "pairs_small.npz"))
    ch = synth.make_chains(len(g["lengths"]), g["lengths"], int(g["d"]), seed=int(g["seed"]), family_size=int(g["family_size"]))
    sel = [int(p) for p in g["msa_sel"]]
    prots = [MA.Protein(f"s{p}", *ch.chain(p), "") for p in sel]
    alignment = {f"s{p}": g["msa_aln"][k] for k, p in enumerate(sel)}
    r, c, t = MA.make_rmsd_coverage_tm_matrix(alignment, prots, superpose_first=False)
    np.testing.assert_allclose(r, g["msa_rmsd"], rtol=1e-9, atol=1e-10)
    assert np.array_equal(c, g["msa_cov"])
    np.testing.assert_allclose(t, g["msa_tm"], rtol=1e-9)
    assert np.all(np.diag(r) == 0) and np.all(np.diag(c) == 1) and np.all(np.diag(t) == 1)


def test_mirror_make_pairwise_matrix(eng):
    g = np.load(os.path.join(G, "pairs_small.npz"))
    ch = synth.make_chains(len(g["lengths"]), g["lengths"], int(g["d"]), seed=int(g["seed"]), family_size=int(g["family_size"]))
    prots = [MA.Protein(f"s{p}", *ch.chain(p), "A" * ch.length(p)) for p in range(ch.n)]
    params = dict(flexible=False, gamma_tensor=7.0, gamma_coords=0.03, verbose=False)
    S64 = MA.MultipleAlignment(prots, precision=engine.FP64).make_pairwise_matrix(params)
    np.testing.assert_allclose(S64, g["score_matrix"], rtol=1e-11)
    S32 = MA.StructureMultiple(prots).make_pairwise_matrix(params)
    np.testing.assert_allclose(S32, g["score_matrix"], rtol=1e-4)
    # flexible=True is the tensor-only matrix (tests/test_gpu_flexible.py has the golden cases)
    F = MA.MultipleAlignment(prots, precision=engine.FP64).make_pairwise_matrix(dict(flexible=True, gamma_tensor=7.0))
    assert F.shape == S64.shape and np.array_equal(F, F.T) and not np.allclose(F, S64)
