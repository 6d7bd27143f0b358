"""Tree-level kernels that run the strips of one problem on concurrent warps (k_fill_s64_mw, k_dtw_fill_mw) and the in-library
composition of the final alignment (crt_msa_compose): same bits as the one-warp kernels / the mirror's own composition, and the
affine DTW against the oracle on shapes that need several rounds of strips."""
import os

import numpy as np
import pytest

from caretta_b200 import engine, synth
from caretta_b200 import multiple_alignment as MA
from caretta_b200 import neighbor_joining as NJ
from oracle import oracle as O

pytestmark = pytest.mark.gpu


class _Env:
    def __init__(self, **kv):
        self.kv, self.old = kv, {}

    def __enter__(self):
        for k, v in self.kv.items():
            self.old[k] = os.environ.get(k)
            os.environ[k] = str(v)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def test_dtw_rounds_of_strips_vs_oracle():
    """m = 1700 is 14 strips of 128 columns: two rounds of the 12-warp kernel; n = 3 is shorter than the skew between warps."""
    rng = np.random.default_rng(5)
    eng = engine.Engine()
    shapes = [(40, 1700), (3, 700), (257, 129), (130, 128), (64, 1537)]
    mats = [rng.random(s) ** 6 for s in shapes]
    for go, ge in [(1.0, 0.01), (0.2, 0.2)]:
        with _Env(CARETTA_B200_DTW_MW=1):
            got = eng.dtw_align_batch(mats, go, ge)
        with _Env(CARETTA_B200_DTW_MW=0):
            one = eng.dtw_align_batch(mats, go, ge)
        for (a1, a2, sc), (b1, b2, sb), S in zip(got, one, mats):
            w1, w2, wsc = O.dtw_align(S, go, ge)
            assert a1.tolist() == w1.tolist() and a2.tolist() == w2.tolist() and sc == wsc
            assert b1.tolist() == w1.tolist() and b2.tolist() == w2.tolist() and sb == wsc
    eng.close()


def _align(ch, tree, prm):
    msa = MA.StructureMultiple.from_chains(ch)
    aln = msa.progressive_align(tree, 1.0, 0.01, 1.0, 0.03, prm, dict(flexible=False))
    last = msa.final_sequences[-1]
    return aln, np.array(last.tensors), np.array(last.coordinates), np.array(msa.final_consensus_weights[-1])


def test_progressive_alignment_same_bits_with_and_without():
    """Nodes of 140 - 420 residues (2 - 4 strips of the float64 kernels and more near the root)."""
    rng = np.random.default_rng(11)
    ch = synth.make_chains(24, list(rng.integers(140, 421, 24)), 10, seed=41, family_size=6)
    prm = dict(MA.DEFAULT_SCORE_PARAMS)
    S = MA.StructureMultiple.from_chains(ch).make_pairwise_matrix(prm)
    tree, _ = NJ.neighbor_joining(S.max() - S)
    with _Env(CARETTA_B200_NODE_MW=1, CARETTA_B200_DTW_MW=1, CARETTA_B200_MSA_COMPOSE=1):
        ref = _align(ch, tree, prm)
    assert len(ref[0]) == 24 and len({len(v) for v in ref[0].values()}) == 1
    for env in (dict(CARETTA_B200_NODE_MW=0, CARETTA_B200_DTW_MW=0, CARETTA_B200_MSA_COMPOSE=0),
                dict(CARETTA_B200_NODE_MW=0, CARETTA_B200_DTW_MW=1, CARETTA_B200_MSA_COMPOSE=1),
                dict(CARETTA_B200_NODE_MW=1, CARETTA_B200_DTW_MW=0, CARETTA_B200_MSA_COMPOSE=0)):
        with _Env(**env):
            got = _align(ch, tree, prm)
        assert list(got[0].keys()) == list(ref[0].keys()), env                 # dictionary order is part of the contract
        for k in ref[0]:
            assert got[0][k].dtype == ref[0][k].dtype == np.int64
            assert np.array_equal(got[0][k], ref[0][k]), (env, k)
        for x, y in zip(got[1:], ref[1:]):
            assert np.array_equal(x, y), env


def test_long_nodes_rounds_of_strips():
    """Three chains of ~1600 residues: the nodes have 13 - 15 strips of 128 columns, i.e. two rounds of the 12-warp kernels, with the
    boundary column of a round going through global memory."""
    ch = synth.make_chains(3, [1600, 1580, 1650], 10, seed=77, family_size=3)
    prm = dict(MA.DEFAULT_SCORE_PARAMS)
    tree = np.array([[0, 3], [1, 3], [3, 2]], dtype=np.int64)     # (0, 1) -> node 3, then (3, 2) -> int-final
    with _Env(CARETTA_B200_NODE_MW=1, CARETTA_B200_DTW_MW=1):
        ref = _align(ch, tree, prm)
    assert len(next(iter(ref[0].values()))) >= 1650
    with _Env(CARETTA_B200_NODE_MW=0, CARETTA_B200_DTW_MW=0):
        got = _align(ch, tree, prm)
    for k in ref[0]:
        assert np.array_equal(got[0][k], ref[0][k]), k
    for x, y in zip(got[1:], ref[1:]):
        assert np.array_equal(x, y)
