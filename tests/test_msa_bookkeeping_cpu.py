"""Host logic of the progressive alignment mirror (caretta_b200.multiple_alignment.MultipleAlignment.progressive_align) on
the CPU: the level batching and the final_alignments / final_sequences bookkeeping, with the oracle standing in for the device
(a stand-in ENGINE for the test only -- the product never imports the oracle), against the reference's own output
(tests/golden/msa.npz, made by oracle/gen_golden_msa.py)."""
import os

import numpy as np
import pytest

from caretta_b200 import multiple_alignment as MA
from caretta_b200 import synth
from oracle import oracle as O

G = os.path.join(os.path.dirname(__file__), "golden")


class OracleEngine:
    """Implements the two engine calls progressive_align makes, node by node on the CPU restatement."""
    calls = 0
    batch_sizes = []

    def progressive_node(self, t1, c1, w1, t2, c2, w2, m1, m2, gt, gc, gw, go, ge):
        OracleEngine.calls += 1
        return O.progressive_node(t1, c1, w1, t2, c2, w2, m1, m2, gt, gc, gw, go, ge)

    def progressive_level(self, children, mults, gt, gc, gw, go, ge):
        OracleEngine.calls += 1
        OracleEngine.batch_sizes.append(len(children))
        return [O.progressive_node(*a, *b, m[0], m[1], gt, gc, gw, go, ge) for (a, b), m in zip(children, mults)]


class OraclePoolEngine(OracleEngine):
    """Adds the device-pool calls (crt_msa_*): sequences live in a list on the 'device', only alignments come back."""
    _msa_generation = 0

    def set_chains(self, coords, tensors, offsets):
        self._chains = (coords, tensors, offsets)

    def msa_begin(self, consensus_weight):
        coords, tensors, off = self._chains
        OraclePoolEngine._msa_generation += 1
        self.pool = [(tensors[off[p]:off[p + 1]], coords[off[p]:off[p + 1]], np.full((off[p + 1] - off[p], 1), float(consensus_weight)))
                     for p in range(len(off) - 1)]
        return len(self.pool)

    def msa_level(self, child1, child2, mults, gt, gc, gw, go, ge):
        OracleEngine.calls += 1
        OracleEngine.batch_sizes.append(len(child1))
        first, out = len(self.pool), []
        for a, b, m in zip(child1, child2, mults):
            a1, a2, tm, cm, wm, sc, st = O.progressive_node(*self.pool[a], *self.pool[b], m[0], m[1], gt, gc, gw, go, ge)
            out.append((a1.astype(np.int32), a2.astype(np.int32), sc, st))
            self.pool.append((tm, cm, wm))
        self.pool_snapshot = list(self.pool)
        return first, out

    def msa_fetch(self, ids):
        return [self.pool[i] for i in ids]


class OraclePoolEngineExt(OraclePoolEngine):
    """The product engine's msa_level_ext: alignments come back as arrays that already end in the -1 sentinel."""

    def msa_level_ext(self, child1, child2, mults, gt, gc, gw, go, ge):
        first, out = self.msa_level(child1, child2, mults, gt, gc, gw, go, ge)
        return first, [(np.append(a1, np.int32(-1)), np.append(a2, np.int32(-1)), sc, st) for a1, a2, sc, st in out]


@pytest.mark.parametrize("name", ["fam8", "ragged12", "mixed40"])
@pytest.mark.parametrize("batch", ["pool", "poolx", "1", "0"])
def test_progressive_align_bookkeeping(monkeypatch, name, batch):
    g = np.load(os.path.join(G, "msa.npz"))
    L = g[f"{name}_lengths"]
    ch = synth.make_chains(len(L), list(L), 10, seed=int(g[f"{name}_seed"]), family_size=int(g[f"{name}_family"]))
    the_engine = OraclePoolEngineExt() if batch == "poolx" else OraclePoolEngine() if batch == "pool" else OracleEngine()
    monkeypatch.setattr(MA, "get_engine", lambda: the_engine)
    monkeypatch.setenv("CARETTA_B200_NODE_BATCH", "0" if batch == "0" else "1")
    OracleEngine.calls, OracleEngine.batch_sizes = 0, []
    msa = MA.StructureMultiple.from_chains(ch)
    aln = msa.progressive_align(g[f"{name}_tree"], 1.0, 0.01, 1.0, 0.03, dict(gamma_tensor=7.0, gamma_coords=0.03), None)
    A = np.array([aln[f"s{p}"] for p in range(ch.n)])
    assert np.array_equal(A, g[f"{name}_aln"])                                     # the reference's final alignment
    assert list(aln) == [str(x) for x in g[f"{name}_fa_members"][g[f"{name}_fa_keys"] == "int-final"]]
    if batch != "0":
        assert OracleEngine.calls < ch.n - 1 and sum(OracleEngine.batch_sizes) == ch.n - 1       # fewer calls than nodes
    else:
        assert OracleEngine.calls == ch.n - 1
    # final_sequences / final_alignments exactly like the reference's attributes (names, order, index arrays)
    assert [s.name for s in msa.final_sequences] == [str(x) for x in g[f"{name}_fs_names"]]
    keys, mem, lens, flat = g[f"{name}_fa_keys"], g[f"{name}_fa_members"], g[f"{name}_fa_lens"], g[f"{name}_fa_flat"]
    got_keys = [(k, m) for k, dct in msa.final_alignments.items() for m in dct]
    assert got_keys == [(str(k), str(m)) for k, m in zip(keys, mem)]
    pos = 0
    for k, m, ln in zip(keys, mem, lens):
        assert np.array_equal(msa.final_alignments[str(k)][str(m)], flat[pos:pos + ln]), (k, m)
        pos += ln
    np.testing.assert_allclose(msa.final_sequences[-1].coordinates, g[f"{name}_final_coords"], rtol=0, atol=1e-9)
    assert np.array_equal(msa.final_consensus_weights[-1], g[f"{name}_final_weights"])


@pytest.mark.parametrize("batch", ["pool", "1"])
def test_alignment_object_pickles_like_the_reference(monkeypatch, batch):
    """--write-class pickles the whole MultipleAlignment (multiple_alignment.py:557-559): the lazily built final_alignments /
    final_sequences / final_consensus_weights travel as the plain dict / lists the reference holds."""
    import copy
    import pickle
    g = np.load(os.path.join(G, "msa.npz"))
    L = g["ragged12_lengths"]
    ch = synth.make_chains(len(L), list(L), 10, seed=int(g["ragged12_seed"]), family_size=int(g["ragged12_family"]))
    the_engine = OraclePoolEngine() if batch == "pool" else OracleEngine()
    monkeypatch.setattr(MA, "get_engine", lambda: the_engine)
    monkeypatch.setenv("CARETTA_B200_NODE_BATCH", "1")
    msa = MA.StructureMultiple.from_chains(ch)
    msa.alignment = msa.progressive_align(g["ragged12_tree"], 1.0, 0.01, 1.0, 0.03, dict(gamma_tensor=7.0, gamma_coords=0.03), None)
    for clone in (pickle.loads(pickle.dumps(msa)), copy.deepcopy(msa)):
        assert type(clone.final_alignments) is dict and type(clone.final_sequences) is list and type(clone.final_consensus_weights) is list
        assert list(clone.final_alignments) == list(msa.final_alignments)
        for k in msa.final_alignments:
            assert list(clone.final_alignments[k]) == list(msa.final_alignments[k])
            assert all(np.array_equal(clone.final_alignments[k][m], msa.final_alignments[k][m]) for m in msa.final_alignments[k])
        assert [s.name for s in clone.final_sequences] == [str(x) for x in g["ragged12_fs_names"]]
        assert np.array_equal(clone.final_sequences[-1].tensors, msa.final_sequences[-1].tensors)
        assert np.array_equal(clone.final_consensus_weights[-1], g["ragged12_final_weights"])
        assert all(np.array_equal(clone.alignment[k], msa.alignment[k]) for k in msa.alignment)


def test_tree_with_forward_reference_is_rejected(monkeypatch):
    monkeypatch.setattr(MA, "get_engine", lambda: OracleEngine())
    ch = synth.make_chains(3, [20, 22, 21], 10, seed=1, family_size=3)
    msa = MA.StructureMultiple.from_chains(ch)
    bad = np.array([[0, 3], [4, 3], [3, 2]], dtype=np.uint64)       # node 4 does not exist when node 3 is built
    with pytest.raises(IndexError):
        msa.progressive_align(bad, 1.0, 0.01, 1.0, 0.03, dict(gamma_tensor=7.0, gamma_coords=0.03), None)


@pytest.mark.parametrize("name", ["fam8", "ragged12", "blocks", "sparse"])
def test_reference_structure_selection(monkeypatch, name):
    """get_reference_structures (host logic over the device-made coverage matrices) against the reference's own output; the
    oracle's coverage matrix stands in for the device call."""
    from tests import consumer_cases as CC
    cons = CC.load()
    monkeypatch.setattr(MA, "make_coverage_gap_distance_matrix", lambda a: O.coverage_gap_matrix(a))
    names = [str(x) for x in cons[f"{name}_pnames"]]
    alignment = {n: cons[f"{name}_aln"][p] for p, n in enumerate(names)}
    for mc in (50, 80):
        first, refs, alone = MA.get_reference_structures(alignment, mc)
        gfirst, grefs, gno = CC.reference_groups(cons, name, mc)
        assert first == names[gfirst]
        assert list(refs.items()) == [(names[k], [names[x] for x in v]) for k, v in grefs.items()]
        assert alone == [names[x] for x in gno]


# ------------------------------------------------------------------------------------------------------------------
# flexible=True and the generic SequenceBase driver: host logic against the reference's flexible golden run
# (tests/golden/flexible.npz, oracle/gen_golden_flexible.py), the oracle standing in for the device calls
# ------------------------------------------------------------------------------------------------------------------
from caretta_b200 import engine as _E  # noqa: E402


def _flex(gc):
    """(flexible_score, flexible_mean) encoded in the gamma_coords sentinels of the node / level / pool calls."""
    return (gc < 0, gc == _E.GC_FLEXIBLE)


class FlexOracleEngine(OraclePoolEngine):
    def progressive_node(self, t1, c1, w1, t2, c2, w2, m1, m2, gt, gc, gw, go, ge):
        fs, fm = _flex(gc)
        a1, a2, tm, cm, wm, sc, st = O.progressive_node(t1, c1, w1, t2, c2, w2, m1, m2, gt, 0.03 if fs else gc, gw, go, ge,
                                                        flexible_score=fs, flexible_mean=fm)
        return a1, a2, tm, (np.zeros((len(a1), 3)) if cm is None else cm), wm, sc, st

    def progressive_level(self, children, mults, gt, gc, gw, go, ge):
        return [self.progressive_node(*a, *b, m[0], m[1], gt, gc, gw, go, ge) for (a, b), m in zip(children, mults)]

    def msa_level(self, child1, child2, mults, gt, gc, gw, go, ge):
        first, out = len(self.pool), []
        for a, b, m in zip(child1, child2, mults):
            a1, a2, tm, cm, wm, sc, st = self.progressive_node(*self.pool[a], *self.pool[b], m[0], m[1], gt, gc, gw, go, ge)
            out.append((a1.astype(np.int32), a2.astype(np.int32), sc, st))
            self.pool.append((tm, cm, wm))
        return first, out

    # the calls of the generic SequenceBase driver
    def sw_align_batch(self, mats, gap=0.0, want_paths=True):
        return [(None, None, O.smith_waterman_score(m, gap), 0) for m in mats]

    def score_matrix(self, t1, c1, t2, c2, gamma_tensor=0.03, gamma_coords=0.03, flexible=False):
        return O.score_matrix(np.asarray(t1, float), c1, np.asarray(t2, float), c2, gamma_tensor, gamma_coords, flexible=flexible), 0

    def dtw_align_batch(self, mats, go, ge):
        return [O.dtw_align(m, go, ge) for m in mats]

    def mean_weights(self, w1, w2, a1, a2):
        return O.mean_weights(w1, w2, a1, a2)


@pytest.mark.parametrize("name", ["fam8", "ragged12", "short5"])
@pytest.mark.parametrize("tag", ["tt", "tf"])
@pytest.mark.parametrize("batch", ["pool", "1", "0"])
def test_flexible_progressive_align_host_logic(monkeypatch, name, tag, batch):
    g = np.load(os.path.join(G, "flexible.npz"))
    L = g[f"{name}_lengths"]
    ch = synth.make_chains(len(L), list(L), 10, seed=int(g[f"{name}_seed"]), family_size=int(g[f"{name}_family"]))
    the_engine = FlexOracleEngine()
    monkeypatch.setattr(MA, "get_engine", lambda: the_engine)
    monkeypatch.setenv("CARETTA_B200_NODE_BATCH", "0" if batch == "0" else "1")
    monkeypatch.setenv("CARETTA_B200_MSA_POOL", "1" if batch == "pool" else "0")
    mean_flex = tag == "tt"
    prots = [MA.Protein(f"s{p}", ch.chain(p)[0], None if mean_flex else ch.chain(p)[1], "A" * ch.length(p)) for p in range(ch.n)]
    msa = MA.MultipleAlignment(prots)
    aln = msa.progressive_align(g[f"{name}_tree"], 1.0, 0.01, 1.0, 0.03, dict(flexible=True, gamma_tensor=7.0, gamma_coords=0.03),
                                dict(flexible=mean_flex))
    assert np.array_equal(np.array([aln[f"s{p}"] for p in range(ch.n)]), g[f"{name}_{tag}_aln"])
    fin = msa.final_sequences[-1]
    assert np.array_equal(fin.tensors, g[f"{name}_{tag}_final_tensors"])
    assert (fin.coordinates is None) == mean_flex
    if not mean_flex:
        np.testing.assert_allclose(fin.coordinates, g[f"{name}_{tag}_final_coords"], rtol=0, atol=1e-10)
    assert np.array_equal(msa.final_consensus_weights[-1], g[f"{name}_{tag}_final_weights"])
    with pytest.raises(ValueError):                                   # coordinate-less nodes cannot be scored rigidly
        MA.MultipleAlignment(prots).progressive_align(g[f"{name}_tree"], 1.0, 0.01, 1.0, 0.03, dict(gamma_tensor=7.0), dict(flexible=True))


class _Feature(MA.SequenceBase):
    """A SequenceBase without shape tensors: the driver calls ITS score / mean functions (host code by definition)."""

    def __init__(self, name, feat):
        self.name, self.feat = name, feat

    def score_function(self, other, gamma=1.0):
        return O.rbf_matrix(self.feat, other.feat, gamma)

    def mean_function(self, other, aln_1, aln_2, name_int):
        out = np.zeros((len(aln_1), self.feat.shape[1]))
        for i, (x, y) in enumerate(zip(aln_1, aln_2)):
            out[i] = other.feat[y] if x == -1 else (self.feat[x] if y == -1 else (self.feat[x] + other.feat[y]) / 2)
        return _Feature(name_int, out)

    def __len__(self):
        return self.feat.shape[0]

    def __str__(self):
        return "X" * len(self)


@pytest.mark.parametrize("name", ["fam8", "ragged12", "short5"])
def test_generic_sequence_base_driver_host_logic(monkeypatch, name):
    """A user-defined tensor-Gaussian SequenceBase is the reference's flexible=True run: same matrix, alignment, consensus, and the
    reference's final_alignments bookkeeping (dictionary order = tree order)."""
    g = np.load(os.path.join(G, "flexible.npz"))
    L = g[f"{name}_lengths"]
    ch = synth.make_chains(len(L), list(L), 10, seed=int(g[f"{name}_seed"]), family_size=int(g[f"{name}_family"]))
    the_engine = FlexOracleEngine()
    monkeypatch.setattr(MA, "get_engine", lambda: the_engine)
    seqs = [_Feature(f"s{p}", ch.chain(p)[0]) for p in range(ch.n)]
    msa = MA.MultipleAlignment(seqs)
    S = msa.make_pairwise_matrix(dict(gamma=7.0))
    assert np.array_equal(S, g[f"{name}_score"]) and np.array_equal(S, S.T) and not np.diag(S).any()
    assert np.array_equal(MA.MultipleAlignment(seqs)._pairwise_matrix_generic(dict(gamma=7.0), batch_bytes=20000), S)   # several device batches
    aln = msa.progressive_align(g[f"{name}_tree"], 1.0, 0.01, 1.0, 0.03, dict(gamma=7.0), None)
    assert np.array_equal(np.array([aln[f"s{p}"] for p in range(ch.n)]), g[f"{name}_tt_aln"])
    assert np.array_equal(msa.final_sequences[-1].feat, g[f"{name}_tt_final_tensors"])
    assert np.array_equal(msa.final_consensus_weights[-1], g[f"{name}_tt_final_weights"])
    assert list(msa.final_alignments)[:ch.n] == [f"s{p}" for p in range(ch.n)] and list(msa.final_alignments)[-1] == "int-final"
    assert list(msa.final_alignments["int-final"]) == list(aln)


def test_msa_level_ext_sentinels_without_a_device():
    """Engine.msa_level_ext hands out views of the packed alignment arrays that end in the -1 sentinel.  A node that fills its whole
    capacity (no column pairs two residues) has no room for it inside its own slots: it must get a copy, and its neighbour's first
    entry must stay intact.  The C call is replaced by a stand-in that writes through the pointers like crt_msa_level does."""
    import ctypes as C
    eng = object.__new__(_E.Engine)
    eng.h = None
    eng._msa_lengths = [2, 3, 1, 1, 4, 2]                      # pool: six sequences
    plan = {0: ([0, 1, -1], [0, 1, 2]),                        # children (0, 1): 3 of 5 slots used
            1: ([0, -1], [-1, 0]),                             # children (2, 3): 2 of 2 slots used -> no room for the sentinel
            2: ([0, 1, 2, 3, -1, -1], [-1, -1, -1, -1, 0, 1])}  # children (4, 5), last node: 6 of 6 slots, spare slot behind it

    class Lib:
        @staticmethod
        def crt_msa_level(h, k, c1, c2, M, gt, gc, gw, go, ge, a1, a2, cap, off, ln, sc, st, first):
            A1, A2 = (C.c_int32 * (cap + 1)).from_address(a1.value), (C.c_int32 * (cap + 1)).from_address(a2.value)
            OFF, LN = (C.c_int64 * (k + 1)).from_address(off.value), (C.c_int32 * k).from_address(ln.value)
            SC, ST = (C.c_double * k).from_address(sc.value), (C.c_int32 * k).from_address(st.value)
            caps, pos = [5, 2, 6], 0
            for q in range(k):
                x, y = plan[q]
                OFF[q] = pos
                for i, (u, v) in enumerate(zip(x, y)):
                    A1[pos + i], A2[pos + i] = u, v
                for i in range(len(x), caps[q]):
                    A1[pos + i], A2[pos + i] = 77, 77       # unused capacity
                LN[q], SC[q], ST[q] = len(x), 1.5 + q, 0
                pos += caps[q]
            OFF[k] = pos
            first._obj.value = 6
            return 0

    eng.lib = Lib()
    first, out = eng.msa_level_ext([0, 2, 4], [1, 3, 5], [(0.25, 0.25)] * 3)
    assert first == 6 and eng._msa_lengths[6:] == [3, 2, 6]
    for q, (x, y) in plan.items():
        assert out[q][0].tolist() == x + [-1] and out[q][1].tolist() == y + [-1] and out[q][0].dtype == np.int32
        assert out[q][2] == 1.5 + q and out[q][3] == 0
