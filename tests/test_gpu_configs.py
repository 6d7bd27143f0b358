"""The contract configurations at FULL size (BASELINE.json configs 4, 5 and the north_star target T), each through the C ABI in
both precisions: the fp32 production matrix (all-vs-all, every pair computed) and the float64 parity mode against the pinned
oracle on >= 500 random pairs -- score / RMSD / TM within 1e-4 (fp32), 1e-11 / identical ncommon (float64) -- plus the
size-independent properties (symmetry, zero diagonal, finite).  Config 5 additionally compares fp32 with float64 on EVERY pair."""
import numpy as np
import pytest

from caretta_b200 import engine, synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    e = engine.Engine()
    yield e
    e.close()


def _sample(n, k, seed):
    rng = np.random.default_rng(seed)
    pi = rng.integers(0, n - 1, size=k)
    pj = np.array([rng.integers(i + 1, n) for i in pi])
    return pi.astype(np.int32), pj.astype(np.int32)


def _check_against_oracle(eng, ch, S, R, T, k, seed):
    pi, pj = _sample(ch.n, k, seed)
    o = O.pairwise_list(ch.coords, ch.tensors, ch.offsets, pi, pj, 7.0, 0.03, 0, extras=True)
    np.testing.assert_allclose(S[pi, pj], o["score"], rtol=1e-4, atol=1e-30)
    np.testing.assert_allclose(R[pi, pj], o["rmsd"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(T[pi, pj], o["tm"], rtol=1e-4, atol=1e-6)
    r64 = eng.pairwise_list(eng.params(precision=engine.FP64), pi, pj)
    np.testing.assert_allclose(r64["score"], o["score"], rtol=1e-11)
    assert np.array_equal(r64["ncommon"], o["ncommon"])
    np.testing.assert_allclose(r64["rmsd"], o["rmsd"], rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(r64["tm"], o["tm"], rtol=1e-8, atol=1e-12)


def _properties(S, R, T):
    assert np.array_equal(S, S.T) and np.all(np.diag(S) == 0) and np.all(np.isfinite(S))
    assert np.array_equal(R, R.T) and np.all(np.diag(R) == 0)
    assert np.array_equal(T, T.T) and np.all(np.diag(T) == 1)


def test_target_T_full_size(eng):
    """north_star target: 5000 chains x 300 residues, 12 497 500 pairs."""
    ch = synth.config("T")
    eng.set_chains(ch.coords, ch.tensors, ch.offsets)
    S, R, T = eng.pairwise_all(eng.params(precision=engine.FP32), want_rmsd_tm=True)
    _properties(S, R, T)
    assert eng.last_rerun()[0] > 0
    _check_against_oracle(eng, ch, S, R, T, 600, 61)


def test_c4_full_size_mixed_lengths(eng):
    """BASELINE config 4: 5000 chains, lengths 50-1000 mixed (every columns-per-lane variant, single and multi strip units,
    the cost-ordered schedule of 12 497 500 pairs)."""
    ch = synth.config("C4")
    eng.set_chains(ch.coords, ch.tensors, ch.offsets)
    S, R, T = eng.pairwise_all(eng.params(precision=engine.FP32), want_rmsd_tm=True)
    _properties(S, R, T)
    _check_against_oracle(eng, ch, S, R, T, 500, 41)


def test_c5_full_size_every_pair(eng):
    """BASELINE config 5: 500 chains x 1500 residues (5 strips per unit, traceback far beyond shared memory).  fp32 against the
    float64 mode on all 124 750 pairs.  Without the tie detection 737 pairs (0.59 %) missed the 1e-4 bound (round 1).  With it the
    pairs that remain outside are alternative alignments whose stage-1 scores agree to the resolution of fp32 score sums (~1e-6
    relative: the exponent of a score is rounded at 2^-22 of ~20): measured 3 pairs (2.4e-5 of the pairs); the bound asserted
    here is 1e-4 of the pairs, and that every such pair's fp32 alignment is as good as the reference's to the accuracy of an fp32
    score sum (stage-1 scores of ~31 within 1e-5; measured 3e-6)."""
    ch = synth.config("C5")
    eng.set_chains(ch.coords, ch.tensors, ch.offsets)
    S32, R32, T32 = eng.pairwise_all(eng.params(precision=engine.FP32), want_rmsd_tm=True)
    n_rerun = eng.last_rerun()[0]
    S64, R64, T64 = eng.pairwise_all(eng.params(precision=engine.FP64), want_rmsd_tm=True)
    _properties(S32, R32, T32)
    pi, pj = np.triu_indices(ch.n, 1)
    rel = np.abs(S32[pi, pj] - S64[pi, pj]) / S64[pi, pj]
    out = rel > 1e-4
    assert out.mean() <= 1e-4, (int(out.sum()), n_rerun)
    if out.any():
        # the stage-1 (tensor) Smith-Waterman score of both alignments: flexible mode returns exactly that score
        q = np.nonzero(out)[0]
        f32 = eng.pairwise_list(eng.params(precision=engine.FP32, flexible=True), pi[q], pj[q])["score"]
        f64 = eng.pairwise_list(eng.params(precision=engine.FP64, flexible=True), pi[q], pj[q])["score"]
        np.testing.assert_allclose(f32, f64, rtol=1e-5)
    _check_against_oracle(eng, ch, S32, R32, T32, 500, 51)
