"""GPU parity of the alignment consumers through the C ABI: coverage/gap matrix, reference selection,
the superpose family, the '%.4f' matrix writer and the FASTA writer -- against the golden vectors produced by the unmodified
reference (tests/golden/consumers.npz) and against the oracle on larger seeded inputs."""
import os

import numpy as np
import pytest

from caretta_b200 import engine, synth
from caretta_b200 import multiple_alignment as MA
from oracle import oracle as O
from tests import consumer_cases as CC

pytestmark = pytest.mark.gpu
TOL = dict(rtol=0, atol=1e-9)          # downstream of a 3x3 SVD (reference: LAPACK gesdd + BLAS; here: Jacobi)


@pytest.fixture(scope="module")
def cons():
    return CC.load()


@pytest.fixture(scope="module")
def eng():
    e = MA.get_engine()
    yield e


def _proteins(ch, gold, name):
    names = [str(x) for x in gold[f"{name}_pnames"]]
    seqs = [str(x) for x in gold[f"{name}_seqs"]]
    return [MA.Protein(names[p], ch.chain(p)[0].copy(), ch.chain(p)[1].copy(), seqs[p]) for p in range(ch.n)]


def _alignment(gold, name):
    names = [str(x) for x in gold[f"{name}_pnames"]]
    return {n: gold[f"{name}_aln"][p] for p, n in enumerate(names)}


@pytest.mark.parametrize("name", CC.CASES)
def test_coverage_gap_and_reference_structures(cons, eng, name):
    aln = cons[f"{name}_aln"]
    dist, al = MA.make_coverage_gap_distance_matrix(aln)
    assert dist.dtype == np.float64 and al.dtype == np.int32
    assert np.array_equal(dist, cons[f"{name}_cg_distance"]) and np.array_equal(al, cons[f"{name}_cg_aligning"])      # bit-exact
    names = [str(x) for x in cons[f"{name}_pnames"]]
    for mc in (50, 80):
        first, refs, no_al = MA.get_reference_structures(_alignment(cons, name), mc)
        gfirst, grefs, gno = CC.reference_groups(cons, name, mc)
        assert first == names[gfirst]
        assert list(refs.items()) == [(names[k], [names[x] for x in v]) for k, v in grefs.items()]
        assert no_al == [names[x] for x in gno]


@pytest.mark.parametrize("name", CC.CASES)
def test_superpose_family_matches_reference(cons, eng, name, capsys):
    ch = CC.chains_of(name, cons)
    alignment = _alignment(cons, name)
    names = list(alignment)
    ref = names[int(cons[f"{name}_reference"])]
    cat = lambda P: np.concatenate([p.coordinates for p in P])      # noqa: E731
    if len(cons[f"{name}_core"]):
        np.testing.assert_allclose(cat(MA.superpose_core(alignment, _proteins(ch, cons, name), ref)), cons[f"{name}_sup_core"], **TOL)
        other = names[(names.index(ref) + 1) % len(names)]
        np.testing.assert_allclose(cat(MA.superpose_core(alignment, _proteins(ch, cons, name), other)), cons[f"{name}_sup_core_other"], **TOL)
        np.testing.assert_allclose(cat(MA.superpose_core(alignment, _proteins(ch, cons, name), ref, core_indices=cons[f"{name}_core"])),
                                   cons[f"{name}_sup_core"], **TOL)
    else:
        with pytest.raises(engine.CrtError):
            MA.superpose_core(alignment, _proteins(ch, cons, name), ref)
    if len(cons[f"{name}_sup_reference"]):
        np.testing.assert_allclose(cat(MA.superpose_reference(alignment, _proteins(ch, cons, name), ref)), cons[f"{name}_sup_reference"], **TOL)
    else:
        with pytest.raises(AssertionError):
            MA.superpose_reference(alignment, _proteins(ch, cons, name), ref)
    if len(cons[f"{name}_sup_auto"]):
        P = _proteins(ch, cons, name)
        Q = MA.superpose(alignment, P)
        assert Q is P and f"Core indices {len(cons[f'{name}_core'])}" in capsys.readouterr().out
        np.testing.assert_allclose(cat(Q), cons[f"{name}_sup_auto"], **TOL)
    for mc in (50, 80):
        if len(cons[f"{name}_suprefs{mc}"]):
            np.testing.assert_allclose(cat(MA.superpose_references(alignment, _proteins(ch, cons, name), mc)), cons[f"{name}_suprefs{mc}"], **TOL)


def test_superpose_outputs_are_consistent(cons, eng):
    """rot / tran returned by the ABI reproduce the coordinates; untouched chains keep theirs bit for bit."""
    name = "ragged12"
    ch = CC.chains_of(name, cons)
    eng.set_chains(ch.coords, ch.tensors, ch.offsets)
    aln = cons[f"{name}_aln"]
    res = eng.superpose(aln, engine.SUP_REFERENCE)
    assert res["mode"] == engine.SUP_REFERENCE and res["reference"] == int(cons[f"{name}_reference"])
    for p in range(ch.n):
        c = ch.chain(p)[1]
        np.testing.assert_allclose(c @ res["rot"][p] + res["tran"][p], res["coords"][ch.offsets[p]:ch.offsets[p + 1]], rtol=0, atol=1e-10)
        R = res["rot"][p]
        np.testing.assert_allclose(R @ R.T, np.eye(3), atol=1e-12)
        assert abs(np.linalg.det(R) - 1) < 1e-12
    # one pair only: everything else is copied through unchanged
    r2 = eng.superpose_pairs(aln, [0], [3], [0, 1])
    keep = np.ones(len(ch.coords), bool)
    keep[ch.offsets[3]:ch.offsets[4]] = False
    assert np.array_equal(r2["coords"][keep], ch.coords[keep])
    p1, p2 = O.common_positions(aln[0], aln[3])
    assert int(r2["ncommon"][0]) == len(p1)
    R, t = O.kabsch(np.ascontiguousarray(ch.chain(0)[1][p1]), np.ascontiguousarray(ch.chain(3)[1][p2]))
    np.testing.assert_allclose(r2["rot"][0], R, atol=1e-10)
    np.testing.assert_allclose(r2["tran"][0], t, atol=1e-8)
    with pytest.raises(engine.CrtError):                     # a member that is also a reference inside one batch
        eng.superpose_pairs(aln, [0, 3], [3, 5], [0, 2])
    with pytest.raises(engine.CrtError):                     # index beyond the chain
        bad = aln.copy()
        bad[2, np.argmax(bad[2])] = 10 ** 6
        eng.superpose(bad)


def test_superpose_large_against_oracle(eng):
    """300 chains x ~200 residues with a synthetic gappy alignment: the three modes against the oracle."""
    rng = np.random.default_rng(11)
    n, A = 300, 260
    keep = rng.random((n, A)) < 0.8
    keep[:, :40] = True
    aln = -np.ones((n, A), np.int64)
    lengths = keep.sum(axis=1)
    for p in range(n):
        aln[p, keep[p]] = np.arange(lengths[p])
    ch = synth.make_chains(n, list(lengths), 10, seed=12, family_size=20)
    coords = [ch.chain(p)[1] for p in range(n)]
    eng.set_chains(ch.coords, ch.tensors, ch.offsets)
    res = eng.superpose(aln, engine.SUP_CORE)
    want, _, _ = O.superpose_core(aln, coords, res["reference"])
    assert res["n_core"] == len(O.core_columns(aln)) >= 40 and res["reference"] == O._reference_index(aln)
    np.testing.assert_allclose(res["coords"], np.concatenate(want), **TOL)
    res = eng.superpose(aln, engine.SUP_REFERENCE)
    np.testing.assert_allclose(res["coords"], np.concatenate(O.superpose_reference(aln, coords, res["reference"])[0]), **TOL)
    d, a = eng.coverage_gap_matrix(aln)
    dw, aw = O.coverage_gap_matrix(aln)
    assert np.array_equal(d, dw) and np.array_equal(a, aw)


def test_rmsd_matrix_superpose_first(cons, eng):
    """make_rmsd_coverage_tm_matrix(superpose_first=True): superpose() then per-pair RMSD / TM in the common frame."""
    name = "fam8"
    ch = CC.chains_of(name, cons)
    alignment = _alignment(cons, name)
    aln = cons[f"{name}_aln"]
    r, c, t = MA.make_rmsd_coverage_tm_matrix(alignment, _proteins(ch, cons, name), superpose_first=True)
    sup = cons[f"{name}_sup_auto"]
    n = ch.n
    for i in range(n):
        for j in range(i + 1, n):
            p1, p2 = O.common_positions(aln[i], aln[j])
            x = np.ascontiguousarray(sup[ch.offsets[i]:ch.offsets[i + 1]][p1])
            y = np.ascontiguousarray(sup[ch.offsets[j]:ch.offsets[j + 1]][p2])
            assert abs(r[i, j] - O.rmsd(x, y)) < 1e-9 and r[j, i] == r[i, j]
            assert abs(t[i, j] - O.tm_score(x, y, ch.lengths[i], ch.lengths[j])) < 1e-9
            assert c[i, j] == len(p1) / aln.shape[1]
    assert np.all(np.diag(r) == 0) and np.all(np.diag(c) == 1) and np.all(np.diag(t) == 1)


def test_matrix_text_golden(cons, eng, tmp_path):
    for key in ("special", "random"):
        M = cons[f"{key}_matrix"]
        names = [f"n{i}" for i in range(M.shape[0])] if key == "special" else [f"id{i}/chain{'A' * (i % 4)}" for i in range(40)]
        assert eng.format_matrix(names, M) == cons[f"{key}_txt"].tobytes(), key
    for name in CC.CASES:
        names = [str(x) for x in cons[f"{name}_pnames"]]
        f = tmp_path / f"{name}.txt"
        MA.write_distance_matrix(names, cons[f"{name}_cg_distance"], f)
        assert f.read_bytes() == cons[f"{name}_dist_txt"].tobytes(), name


def test_matrix_text_against_printf(eng):
    """Rounding of '%.4f': random doubles over the whole exponent range, values on and next to decimal ties, 3000 x 3000 at scale."""
    rng = np.random.default_rng(5)
    k = np.arange(0, 20000)
    ties = (2 * k + 1) / 20000.0                                          # x.xxxx5 as closely as a double gets
    near = np.concatenate([ties, np.nextafter(ties, 0), np.nextafter(ties, 1), ties * 1000, -ties, k / 16.0 + 1 / 32.0])
    bits = rng.integers(0, 2 ** 63, size=40000, dtype=np.int64).view(np.float64)
    bits = bits[np.isfinite(bits)]
    wide = np.concatenate([bits, -bits[:1000], rng.normal(size=20000) * 10.0 ** rng.integers(-8, 20, 20000)])
    for vals in (near, wide):
        n = int(np.sqrt(len(vals)))
        M = vals[:n * n].reshape(n, n)
        names = [f"r{i}" for i in range(n)]
        assert eng.format_matrix(names, M) == O.format_matrix(names, M)
    M = rng.random((3000, 3000)) * 100
    names = [f"protein_{i:04d}" for i in range(3000)]
    got = eng.format_matrix(names, M)
    assert got == O.format_matrix(names, M)
    # edge shapes: no columns, one cell, empty
    assert eng.format_matrix(["a", "bb"], np.zeros((2, 0))) == b"2\na \nbb \n"
    assert eng.format_matrix(["a"], np.array([[2.5]])) == b"1\na 2.5000\n"
    assert eng.format_matrix([], np.zeros((0, 0))) == b"0\n"


def test_fasta_golden(cons, eng, tmp_path):
    for name in CC.CASES:
        ch = CC.chains_of(name, cons)
        msa = MA.MultipleAlignment(_proteins(ch, cons, name), alignment=_alignment(cons, name))
        f = tmp_path / f"{name}.fasta"
        msa.write_alignment(f)
        assert f.read_bytes() == cons[f"{name}_fasta"].tobytes(), name
        seqaln = msa.to_sequence_alignment()
        lines = cons[f"{name}_fasta"].tobytes().decode().split("\n")
        assert seqaln == {lines[2 * q][1:]: lines[2 * q + 1] for q in range(ch.n)}
    with pytest.raises(engine.CrtError):
        eng.format_fasta(["a"], ["ACD"], np.array([[0, 1, 2, 3]]))


def test_fast_mode_guide_matrix(cons, eng):
    """make_count_matrix / braycurtis (multiple_alignment.py:128-145): bit-identical to the reference's numba output, and to the
    oracle on a larger random case."""
    off = np.concatenate([[0], np.cumsum(cons["bc_lengths"])])
    res = [cons["bc_indices"][off[p]:off[p + 1]] for p in range(len(cons["bc_lengths"]))]
    counts = MA.make_count_matrix(res, 1024)
    assert counts.dtype == np.float64 and np.array_equal(counts, cons["bc_counts"])
    assert np.array_equal(MA.braycurtis(counts, counts), cons["bc_dist"])
    assert np.array_equal(MA.braycurtis(cons["bc_x"], cons["bc_y"]), cons["bc_xy"])
    rng = np.random.default_rng(17)
    big = [rng.integers(0, 1024, int(rng.integers(50, 400))) for _ in range(301)]
    cb = eng.count_matrix(big, 1024)
    assert np.array_equal(cb, O.count_matrix(big, 1024))
    assert np.array_equal(eng.braycurtis(cb, cb[:77]), O.braycurtis(cb, cb[:77]))
    with pytest.raises(engine.CrtError):
        eng.count_matrix([np.array([5, 1024])], 1024)


def test_consumer_edge_cases(eng, tmp_path):
    # smallest alignments
    d, a = eng.coverage_gap_matrix(np.array([[0]]))
    assert d.tolist() == [[0.0]] and a.tolist() == [[1]]
    d, a = eng.coverage_gap_matrix(np.array([[0, -1, 1], [-1, 0, 1]]))
    assert d.tolist() == [[0.0, 0.5], [0.5, 0.0]] and a.tolist() == [[2, 1], [1, 2]]
    with pytest.raises(engine.CrtError):                                   # a protein without any residue: the reference divides by zero
        eng.coverage_gap_matrix(np.array([[0, 1], [-1, -1]]))
    with pytest.raises(engine.CrtError):
        eng.coverage_gap_matrix(np.array([[0, -2]]))
    # two proteins, exactly 4 common columns (the reference's assert needs > 3), columns beyond one chain's end are errors
    ch = synth.make_chains(2, [6, 5], 10, seed=9, family_size=2)
    eng.set_chains(ch.coords, ch.tensors, ch.offsets)
    aln = np.array([[0, 1, 2, 3, 4, 5, -1], [0, 1, -1, 2, 3, -1, 4]])
    res = eng.superpose(aln, engine.SUP_REFERENCE)
    assert res["reference"] == 0 and res["ncommon"].tolist() == [6, 4] and res["n_core"] == 4
    want = O.superpose_reference(aln, [ch.chain(0)[1], ch.chain(1)[1]], 0)[0]
    np.testing.assert_allclose(res["coords"], np.concatenate(want), **TOL)
    res = eng.superpose(aln, engine.SUP_AUTO)                              # 4 core columns of 7 >= 7 // 2 -> core branch
    assert res["mode"] == engine.SUP_CORE
    np.testing.assert_allclose(res["coords"], np.concatenate(O.superpose_core(aln, [ch.chain(0)[1], ch.chain(1)[1]], 0)[0]), **TOL)
    with pytest.raises(engine.CrtError):
        eng.superpose(np.array([[0, 1, 2, 3, 4, 6], [0, 1, 2, 3, 4, -1]]))
    with pytest.raises(engine.CrtError):                                   # a supplied core column that holds a gap
        eng.superpose(aln, engine.SUP_CORE, 0, core_columns=[0, 1, 2])
    none = eng.superpose_pairs(aln, [], [], [0])                           # no pairs: coordinates pass through
    assert np.array_equal(none["coords"], ch.coords)
    few = eng.superpose_pairs(np.array([[0, 1, 2, -1, -1, -1, -1], [-1, 0, 1, 2, 3, 4, -1]]), [0], [1], [0, 1])
    assert few["ncommon"].tolist() == [2] and np.array_equal(few["coords"], ch.coords)          # <= 3 common: skipped
    # text: UTF-8 names, one row, gaps only
    names = ["protéine/α", "b"]
    M = np.array([[0.0, 1.23456], [1.23456, 0.0]])
    f = tmp_path / "m.txt"
    MA.write_distance_matrix(names, M, f)
    assert f.read_bytes() == (f"2\n{names[0]} 0.0000 1.2346\n{names[1]} 1.2346 0.0000\n").encode("utf-8")
    msa = MA.MultipleAlignment([MA.Protein("x", np.zeros((3, 10)), np.zeros((3, 3)), "ACD"), MA.Protein("y", np.zeros((2, 10)), np.zeros((2, 3)), "KL")],
                               alignment={"x": np.array([0, -1, 1, 2]), "y": np.array([-1, -1, -1, -1])})
    assert msa.to_sequence_alignment() == {"x": "A-CD", "y": "----"}
    msa.write_alignment(tmp_path / "a.fasta")
    assert (tmp_path / "a.fasta").read_bytes() == b">x\nA-CD\n>y\n----\n"


def test_coordinates_only_chain_set(cons, eng):
    """crt_set_coords: enough for the consumers, refused by the pair path."""
    ch = CC.chains_of("fam8", cons)
    aln = cons["fam8_aln"]
    eng.set_chains(ch.coords, ch.tensors, ch.offsets)
    want = eng.superpose(aln, engine.SUP_CORE)
    r0 = eng.rmsd_cov_tm(aln)
    eng.set_coords(ch.coords, ch.offsets)
    got = eng.superpose(aln, engine.SUP_CORE)
    assert np.array_equal(got["coords"], want["coords"]) and np.array_equal(got["rot"], want["rot"])
    r1 = eng.rmsd_cov_tm(aln)
    assert all(np.array_equal(a, b) for a, b in zip(r0[:3], r1[:3]))
    with pytest.raises(engine.CrtError):
        eng.pairwise_all(eng.params())
    eng.set_chains(ch.coords, ch.tensors, ch.offsets)
    assert eng.pairwise_all(eng.params()).shape == (ch.n, ch.n)


def test_superpose_takes_the_first_alignment_key_among_ties(capsys):
    """superpose() (multiple_alignment.py:896-910): the reference structure is the first key of `alignment` -- guide-tree order,
    not the order of `proteins` -- among those with the most aligned residues (sorted() is stable), and the core columns are
    gap-free over every entry of the alignment, also over entries whose protein is not in `proteins`."""
    rng = np.random.default_rng(9)
    n = 40
    base = np.cumsum(rng.normal(size=(n, 3)), axis=0) * 3.8
    def prot(name, noise):
        return MA.Protein(name, rng.normal(size=(n, 10)), base @ _rot(rng) + rng.normal(size=3) * 10 + rng.normal(size=(n, 3)) * noise, "A" * n)
    def _rot(r):
        q, _ = np.linalg.qr(r.normal(size=(3, 3)))
        return q * np.sign(np.linalg.det(q))
    full = np.arange(n, dtype=np.int64)
    proteins = [prot("a", 0.3), prot("b", 0.3), prot("c", 0.3)]
    alignment = {"c": full.copy(), "a": full.copy(), "b": full.copy()}             # all tie on length: the reference is "c"
    want = MA.superpose_core(alignment, [MA.Protein(p.name, p.tensors, p.coordinates.copy(), p.sequence) for p in proteins], "c")
    got = MA.superpose(alignment, [MA.Protein(p.name, p.tensors, p.coordinates.copy(), p.sequence) for p in proteins])
    assert "Core indices 40" in capsys.readouterr().out
    for g, w in zip(got, want):
        np.testing.assert_array_equal(g.coordinates, w.coordinates)
    other = MA.superpose_core(alignment, [MA.Protein(p.name, p.tensors, p.coordinates.copy(), p.sequence) for p in proteins], "a")
    assert not np.allclose(other[1].coordinates, want[1].coordinates)
    # an alignment entry without a protein still counts for the core columns: its gaps shrink the core to 30 columns
    alignment["ghost"] = np.concatenate([full[:30], -np.ones(10, np.int64)])
    MA.superpose(alignment, [MA.Protein(p.name, p.tensors, p.coordinates.copy(), p.sequence) for p in proteins])
    assert "Core indices 30" in capsys.readouterr().out
