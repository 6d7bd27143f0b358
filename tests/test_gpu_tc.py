"""Stage 1 of the fp32 mode on the tensor-core kernel (k_fill1_tc + k_trace_tc, crt_fill_tc.cuh; CARETTA_B200_TC=1, experimental
and off by default).  What is pinned here: on BASELINE config 2 the tensor-core path takes the reference's alignment on every
one of the 19 900 pairs (golden vectors from the unmodified reference), it really runs (crt_last_tc_pairs), rounds mix with
left-over systolic units without touching the result slots, and the raw decisions of the two fp32 stage-1 kernels agree on all
but a handful of pairs of a 400 x 300 set.  What is NOT claimed: at full size (C3, C5) one to three pairs per run take an unmarked
different path (profiles/r02_tensor_core_stage1.md), which is why the kernel is not the default."""
import os

import numpy as np
import pytest

from caretta_b200 import engine, synth

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def eng():
    e = engine.Engine(0)
    yield e
    e.close()


@pytest.fixture()
def tc_on(monkeypatch):
    monkeypatch.setenv("CARETTA_B200_TC", "1")


def _cols(a1, a2):
    return {(int(x), int(y)) for x, y in zip(a1, a2) if x >= 0 and y >= 0}


def test_tc_c2_takes_the_reference_alignment_on_every_pair(eng, tc_on):
    g = np.load(os.path.join(G, "c2_full.npz"))
    ch = synth.config("C2")
    eng.set_chains(ch.coords, ch.tensors, ch.offsets)
    pi, pj = np.triu_indices(ch.n, 1)
    res = eng.pairwise_list(eng.params(precision=engine.FP32), pi, pj, want_paths=True)
    assert eng.last_tc_pairs() > 10000                       # columns with >= 72 partners run in rounds of 128
    off, goff = res["aln_off"], g["aln_off"]
    for q in range(len(pi)):
        assert _cols(res["aln1"][off[q]:off[q + 1]], res["aln2"][off[q]:off[q + 1]]) == \
            _cols(g["aln1"][goff[q]:goff[q + 1]], g["aln2"][goff[q]:goff[q + 1]]), q
    np.testing.assert_allclose(res["score"], g["score"], rtol=1e-4, atol=1e-30)
    np.testing.assert_allclose(res["rmsd"], g["rmsd"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(res["tm"], g["tm"], rtol=1e-4, atol=1e-6)
    assert np.array_equal(res["ncommon"], g["ncommon"])


def test_tc_matrix_equals_the_systolic_matrix_within_tolerance(eng, monkeypatch):
    """Ragged chains (strips of 16..160 columns, half tiles, partners of different lengths in one round, left-over units): the
    all-vs-all matrices of the two stage-1 kernels agree to 1e-4 on all but at most 2 of the 79 800 pairs, and the raw stage-1
    decisions (no float64 re-run) give the same number of matched residues on >= 99.9 % of the pairs."""
    rng = np.random.default_rng(5)
    ch = synth.make_chains(400, rng.integers(20, 400, 400), 10, seed=11, family_size=20)
    eng.set_chains(ch.coords, ch.tensors, ch.offsets)
    prm = eng.params(precision=engine.FP32)
    out = {}
    for tc in ("0", "1"):
        monkeypatch.setenv("CARETTA_B200_TC", tc)
        S, R, T = eng.pairwise_all(prm, want_rmsd_tm=True)
        n_tc = eng.last_tc_pairs()
        monkeypatch.setenv("CARETTA_B200_TIE_RERUN", "0")
        eng.pairwise_shard(prm, 0, 1)
        raw = eng.fetch(eng.shard_size(0, 1))
        monkeypatch.delenv("CARETTA_B200_TIE_RERUN")
        out[tc] = (S, R, T, n_tc, raw)
    assert out["0"][3] == 0 and out["1"][3] > 60000
    pi, pj = np.triu_indices(ch.n, 1)
    s0, s1 = out["0"][0][pi, pj], out["1"][0][pi, pj]
    far = np.abs(s0 - s1) > 1e-4 * np.maximum(s0, 1e-30)
    assert far.sum() <= 2, (int(far.sum()), list(zip(pi[far][:5], pj[far][:5])))
    same_nc = out["0"][4]["ncommon"] == out["1"][4]["ncommon"]
    assert same_nc.mean() >= 0.999, same_nc.mean()
    assert np.array_equal(out["1"][0], out["1"][0].T) and np.all(np.diag(out["1"][0]) == 0)
