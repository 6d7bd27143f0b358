"""Pins the C restatement (oracle/) against vectors produced by the unmodified reference
(oracle/gen_golden.py -> tests/golden/*.npz).  Bit-exact for RBF / SW / DTW / common positions;
1e-10 for anything downstream of the Kabsch SVD (reference: LAPACK gesdd + BLAS gemm)."""
import os

import numpy as np
import pytest

from caretta_b200 import synth
from oracle import oracle as O

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def kat():
    return np.load(os.path.join(G, "kat.npz"))


@pytest.fixture(scope="module")
def small():
    return np.load(os.path.join(G, "pairs_small.npz"))


@pytest.fixture(scope="module")
def small_chains(small):
    ch = synth.make_chains(len(small["lengths"]), small["lengths"], int(small["d"]), seed=int(small["seed"]),
                           family_size=int(small["family_size"]))
    import hashlib
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(ch.coords).tobytes())
    h.update(np.ascontiguousarray(ch.tensors).tobytes())
    assert h.hexdigest() == str(small["input_digest"]), "synthetic generator drifted from the golden inputs"
    return ch


def test_kat_dtw(kat):
    a1, a2, sc, M, B = O.dtw_align(kat["dtw0_S"], 1.0, 0.01, want_matrices=True)
    assert a1.tolist() == kat["dtw0_a1"].tolist() == [0, 1, -1]
    assert a2.tolist() == kat["dtw0_a2"].tolist() == [0, 1, 2]
    assert sc == float(kat["dtw0_sc"]) == 1.0
    # interior cells and the initialised borders both match (the untouched corners are MIN in both)
    assert np.array_equal(M, kat["dtw0_M"])
    assert np.array_equal(B, kat["dtw0_B"])
    for k in range(1, 7):
        S = kat[f"dtw{k}_S"]
        a1, a2, sc = O.dtw_align(S, float(kat[f"dtw{k}_go"]), float(kat[f"dtw{k}_ge"]))
        assert a1.tolist() == kat[f"dtw{k}_a1"].tolist(), k
        assert a2.tolist() == kat[f"dtw{k}_a2"].tolist(), k
        assert sc == float(kat[f"dtw{k}_sc"]), k


def test_kat_sw(kat):
    for k in range(5):
        S = kat[f"sw{k}_S"]
        a1, a2, sc = O.smith_waterman(S, 0.0)
        assert a1.tolist() == kat[f"sw{k}_a1"].tolist(), k
        assert a2.tolist() == kat[f"sw{k}_a2"].tolist(), k
        assert sc == float(kat[f"sw{k}_sc"]), k
        assert O.smith_waterman_score(S, 0.0) == float(kat[f"sw{k}_score_only"]), k
    assert bool(kat["sw_zero_raises"])
    with pytest.raises(ValueError):
        O.smith_waterman(np.zeros((3, 3)), 0.0)
    assert O.smith_waterman_score(np.zeros((3, 3))) == float(kat["sw_zero_score_only"]) == 0.0
    # Appendix B: first row-major maximum, not the last column
    a1, a2, sc = O.smith_waterman(np.full((3, 4), .5))
    assert (a1.tolist(), a2.tolist(), sc) == ([0, 1, 2], [0, 1, 2], 1.5)


def test_kat_tm_rmsd_kabsch_common(kat):
    assert O.tm_score(np.zeros((2, 3)), np.zeros((2, 3)), 2, 2) == float(kat["tm_zero"]) == 1.0
    assert O.tm_score(kat["tm_x"], kat["tm_y"], 40, 23) == float(kat["tm_val"])          # bit-exact, quirks kept
    assert O.rmsd(kat["tm_x"], kat["tm_y"]) == float(kat["rmsd_val"])
    R, t = O.kabsch(kat["tm_x"], kat["tm_y"])
    np.testing.assert_allclose(R, kat["kab_R"], atol=1e-12)
    np.testing.assert_allclose(t, kat["kab_t"], atol=1e-11)
    np.testing.assert_allclose(O.apply_rotran(kat["tm_y"], R, t), kat["kab_applied"], atol=1e-10)
    R, t = O.kabsch(kat["tm_x"], kat["kab_mirror_y"])
    np.testing.assert_allclose(R, kat["kab_mirror_R"], atol=1e-12)
    assert abs(np.linalg.det(R) - 1.0) < 1e-12      # reflection fix gives a proper rotation
    p1, p2 = O.common_positions(kat["cp_a"], kat["cp_b"])
    assert p1.tolist() == kat["cp_p1"].tolist() and p2.tolist() == kat["cp_p2"].tolist()


def test_rbf_bit_exact(small, small_chains):
    ch = small_chains
    for key in small.files:
        if not key.startswith("ST_"):
            continue
        _, i, j = key.split("_")
        i, j = int(i), int(j)
        ti, ci = ch.chain(i)
        tj, cj = ch.chain(j)
        assert np.array_equal(O.rbf_matrix(ti, tj, 7.0), small[key]), key     # every bit of every cell


def test_pairs_small_paths_scores(small, small_chains):
    ch = small_chains
    off = small["aln_off"]
    n_skip = 0
    for q, (i, j) in enumerate(zip(small["pi"], small["pj"])):
        ti, ci = ch.chain(int(i))
        tj, cj = ch.chain(int(j))
        r = O.pair(ti, ci, tj, cj)
        assert r["aln1"].tolist() == small["aln1"][off[q]:off[q + 1]].tolist(), (i, j)
        assert r["aln2"].tolist() == small["aln2"][off[q]:off[q + 1]].tolist(), (i, j)
        assert r["score1"] == small["score1"][q]
        assert r["ncommon"] == small["ncommon"][q]
        n_skip += r["status"] & 1
        assert (r["status"] & 1) == (small["ncommon"][q] <= 3)
        np.testing.assert_allclose(r["R"], small["R"][q], atol=1e-10)
        np.testing.assert_allclose(r["score"], small["score"][q], rtol=1e-11)
        np.testing.assert_allclose(r["rmsd"], small["rmsd"][q], rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(r["tm"], small["tm"][q], rtol=1e-9, atol=1e-12)
        key = f"SC_{i}_{j}"
        if key in small.files:       # stage-2 matrix: downstream of the SVD, so tolerance
            a1 = small["aln1"][off[q]:off[q + 1]].astype(np.int64)
            a2 = small["aln2"][off[q]:off[q + 1]].astype(np.int64)
            p1, p2 = O.common_positions(a1, a2)
            if len(p1) > 3:
                w1, w2, _ = O.superpose_with_subset(ci, cj, ci[p1], cj[p2])
            else:
                w1, w2 = ci, cj
            np.testing.assert_allclose(O.rbf_matrix(w1, w2, 0.03), small[key], rtol=1e-9, atol=1e-300)
    assert n_skip >= 10          # the length-3/4/5 chains exercise the "<= 3 common positions" branch
    S = O.pairwise_all(ch.coords, ch.tensors, ch.offsets)
    np.testing.assert_allclose(S, small["score_matrix"], rtol=1e-11)
    assert np.all(np.diag(S) == 0) and np.array_equal(S, S.T)


def test_rmsd_cov_tm_matrix(small, small_chains):
    sel = small["msa_sel"]
    ch = small_chains
    coords = np.concatenate([ch.chain(int(p))[1] for p in sel])
    off = np.zeros(len(sel) + 1, np.int64)
    off[1:] = np.cumsum([ch.length(int(p)) for p in sel])
    r, c, t, bad = O.rmsd_cov_tm(small["msa_aln"], coords, off)
    assert bad == 0
    np.testing.assert_allclose(r, small["msa_rmsd"], rtol=1e-9, atol=1e-10)
    assert np.array_equal(c, small["msa_cov"])
    np.testing.assert_allclose(t, small["msa_tm"], rtol=1e-9)


def test_dtw_golden():
    g = np.load(os.path.join(G, "dtw.npz"))
    n = 0
    for q in range(int(g["n_cases"])):
        S = g[f"c{q}_S"]
        for go, ge in [(1.0, 0.01), (0.5, 0.1), (0.0, 0.0)]:
            a1, a2, sc = O.dtw_align(S, go, ge)
            key = f"c{q}_{go}_{ge}"
            assert a1.tolist() == g[f"{key}_a1"].tolist(), key
            assert a2.tolist() == g[f"{key}_a2"].tolist(), key
            assert sc == float(g[f"{key}_sc"]), key
            n += 1
    for q in range(6):
        S = g[f"r{q}_S"]
        a1, a2, sc = O.dtw_align(S, 1.0, 0.01)
        assert a1.tolist() == g[f"r{q}_a1"].tolist() and a2.tolist() == g[f"r{q}_a2"].tolist()
        assert sc == float(g[f"r{q}_sc"])
        a1, a2, sc = O.smith_waterman(S, 0.0)
        assert a1.tolist() == g[f"r{q}_sw_a1"].tolist() and a2.tolist() == g[f"r{q}_sw_a2"].tolist()
        assert sc == float(g[f"r{q}_sw_sc"])
    assert n == 24


def test_c1_test_data():
    g = np.load(os.path.join(G, "c1_test_data.npz"))
    names = [str(n) for n in g["names"]]
    coords = np.concatenate([g[f"ca_{n}"] for n in names])
    tens = np.concatenate([g[f"tensors_{n}"] for n in names])
    off = np.zeros(4, np.int64)
    off[1:] = np.cumsum([len(g[f"ca_{n}"]) for n in names])
    assert off.tolist() == [0, 85, 164, 244]
    S = O.pairwise_all(coords, tens, off)
    np.testing.assert_allclose(S, g["score_matrix"], rtol=1e-11)


def test_c2_full_paths_bit_exact():
    """All 19 900 pairs of config C2 (200 x 80): stage-1 paths identical to the reference's, scores to 1e-11."""
    g = np.load(os.path.join(G, "c2_full.npz"))
    ch = synth.config("C2")
    N = ch.n
    pi, pj = np.triu_indices(N, 1)
    order = np.lexsort((pj, pi))
    pi, pj = pi[order].astype(np.int32), pj[order].astype(np.int32)
    res = O.pairwise_list(ch.coords, ch.tensors, ch.offsets, pi, pj)
    np.testing.assert_allclose(res["score"], g["score"], rtol=1e-11)
    assert np.array_equal(res["ncommon"], g["ncommon"])
    np.testing.assert_allclose(res["rmsd"], g["rmsd"], rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(res["tm"], g["tm"], rtol=1e-8, atol=1e-12)
    off = g["aln_off"]
    rng = np.random.default_rng(0)
    for q in rng.choice(len(pi), 600, replace=False):      # path spot-check (ncommon above covers all pairs)
        r = O.pair(*ch.chain(int(pi[q])), *ch.chain(int(pj[q])))
        assert r["aln1"].tolist() == g["aln1"][off[q]:off[q + 1]].tolist()
        assert r["aln2"].tolist() == g["aln2"][off[q]:off[q + 1]].tolist()


def test_nj_golden():
    """neighbor_joining.py:17-157 restated in C (crt_o_neighbor_joining) against the reference's own output."""
    g = np.load(os.path.join(G, "nj.npz"))
    for name in [str(n) for n in g["names"]]:
        tree, bl = O.neighbor_joining(g[f"{name}_D"])
        assert np.array_equal(tree, g[f"{name}_tree"]), name
        assert np.array_equal(bl, g[f"{name}_bl"]), name
    with pytest.raises(IndexError):
        O.neighbor_joining(np.zeros((2, 2)))


def test_progressive_align_golden():
    """oracle.progressive_align (multiple_alignment.py:172-253 restated on the C primitives) against the reference's
    multiple_align output: identical final alignment, consensus node within 1e-11."""
    g = np.load(os.path.join(G, "msa.npz"))
    for name in ("fam8", "ragged12", "mixed40"):
        L = g[f"{name}_lengths"]
        ch = synth.make_chains(len(L), list(L), 10, seed=int(g[f"{name}_seed"]), family_size=int(g[f"{name}_family"]))
        S = O.pairwise_all(ch.coords, ch.tensors, ch.offsets)
        np.testing.assert_allclose(S, g[f"{name}_score"], rtol=1e-12)
        tree, bl = O.neighbor_joining(np.max(g[f"{name}_score"]) - g[f"{name}_score"])
        assert np.array_equal(tree, g[f"{name}_tree"]) and np.array_equal(bl, g[f"{name}_bl"])
        aln, fs, fw = O.progressive_align([(f"s{p}",) + ch.chain(p) for p in range(ch.n)], tree)
        A = np.array([aln[f"s{p}"] for p in range(ch.n)])
        assert np.array_equal(A, g[f"{name}_aln"])
        np.testing.assert_allclose(fs[-1][2], g[f"{name}_final_coords"], rtol=0, atol=1e-11)
        assert np.array_equal(fw[-1], g[f"{name}_final_weights"])


def test_flexible_golden():
    """flexible=True (tensor-only score matrices, multiple_alignment.py:323-326; coordinate-less nodes, :359-360): the oracle's
    restatement against the reference's outputs (oracle/gen_golden_flexible.py) -- score matrix bit-exact, identical alignments
    for mean_function flexible=True ("tt") and False ("tf")."""
    g = np.load(os.path.join(G, "flexible.npz"))
    for name in ("fam8", "ragged12", "short5"):
        L = g[f"{name}_lengths"]
        ch = synth.make_chains(len(L), list(L), 10, seed=int(g[f"{name}_seed"]), family_size=int(g[f"{name}_family"]))
        S = O.pairwise_all_flexible(ch.tensors, ch.offsets, 7.0)
        assert np.array_equal(S, g[f"{name}_score"]), name
        tree, _ = O.neighbor_joining(np.max(S) - S)
        assert np.array_equal(tree, g[f"{name}_tree"])
        for tag, mean_flex in (("tt", True), ("tf", False)):
            aln, fs, fw = O.progressive_align([(f"s{p}",) + ch.chain(p) for p in range(ch.n)], tree, flexible_score=True,
                                              flexible_mean=mean_flex)
            A = np.array([aln[f"s{p}"] for p in range(ch.n)])
            assert np.array_equal(A, g[f"{name}_{tag}_aln"]), (name, tag)
            assert np.array_equal(fs[-1][1], g[f"{name}_{tag}_final_tensors"])
            assert np.array_equal(fw[-1], g[f"{name}_{tag}_final_weights"])
            if mean_flex:
                assert fs[-1][2] is None
            else:
                np.testing.assert_allclose(fs[-1][2], g[f"{name}_{tag}_final_coords"], rtol=0, atol=1e-10)
    L = g["two_lengths"]
    ch = synth.make_chains(2, list(L), 10, seed=int(g["two_seed"]), family_size=int(g["two_family"]))
    a1, a2, _ = O.dtw_align(O.score_matrix(*ch.chain(0), *ch.chain(1), 7.0, 0.03, flexible=True), 1.0, 0.01)
    assert np.array_equal(np.array([a1, a2]), g["two_tt_aln"]) and np.array_equal(g["two_tt_aln"], g["two_tf_aln"])


def test_protein_methods_golden():
    """Protein.score_function / mean_function / get_mean_weights (multiple_alignment.py:321-383, :73-82) called directly on the
    reference: the oracle's score_matrix / mean_function / mean_weights against its outputs, both flexible settings."""
    g = np.load(os.path.join(G, "flexible.npz"))
    for name in ("fn_a", "fn_short", "fn_b"):
        lens, seed = [int(x) for x in g[f"{name}_lengths"]], int(g[f"{name}_seed"])
        ch = synth.make_chains(2, lens, 10, seed=seed, family_size=2)
        (t1, c1), (t2, c2) = ch.chain(0), ch.chain(1)
        rng = np.random.default_rng(seed)
        w1, w2 = rng.integers(1, 5, (lens[0], 1)).astype(np.float64), rng.integers(1, 4, (lens[1], 1)).astype(np.float64)
        for tag, flex in (("rigid", False), ("flex", True)):
            S = O.score_matrix(t1, c1, t2, c2, 7.0, 0.03, flexible=flex)
            if flex:
                assert np.array_equal(S, g[f"{name}_{tag}_S"])
            else:
                np.testing.assert_allclose(S, g[f"{name}_{tag}_S"], rtol=1e-10, atol=1e-300)
            a1, a2, _ = O.dtw_align(g[f"{name}_{tag}_S"], 1.0, 0.01)
            assert np.array_equal(np.array([a1, a2]), g[f"{name}_{tag}_aln"])
            tm, cm = O.mean_function(t1, c1, t2, c2, a1, a2, flexible=flex)
            assert np.array_equal(tm, g[f"{name}_{tag}_tensors"])
            if not flex:
                np.testing.assert_allclose(cm, g[f"{name}_{tag}_coords"], rtol=0, atol=1e-10)
            assert np.array_equal(O.mean_weights(w1, w2, a1, a2), g[f"{name}_{tag}_weights"])


# ------------------------------------------------------------------------------------------------------------------
# Consumers of the multiple alignment: the oracle restatement against the reference's outputs
# ------------------------------------------------------------------------------------------------------------------
from tests import consumer_cases as CC  # noqa: E402


@pytest.fixture(scope="module")
def cons():
    return CC.load()


@pytest.mark.parametrize("name", CC.CASES)
def test_consumers_coverage_and_references(cons, name):
    aln = cons[f"{name}_aln"]
    dist, al = O.coverage_gap_matrix(aln)
    assert np.array_equal(dist, cons[f"{name}_cg_distance"])            # integer counts and one division: bit-exact
    assert np.array_equal(al, cons[f"{name}_cg_aligning"])
    assert O._reference_index(aln) == int(cons[f"{name}_reference"])
    assert np.array_equal(O.core_columns(aln), cons[f"{name}_core"])
    for mc in (50, 80):
        first, refs, no_al = O.get_reference_structures(aln, mc)
        gfirst, grefs, gno = CC.reference_groups(cons, name, mc)
        assert first == gfirst and list(refs.items()) == list(grefs.items()) and no_al == gno


@pytest.mark.parametrize("name", CC.CASES)
def test_consumers_superposition(cons, name):
    ch = CC.chains_of(name, cons)
    aln = cons[f"{name}_aln"]
    coords = [ch.chain(p)[1] for p in range(ch.n)]
    ref = int(cons[f"{name}_reference"])
    tol = dict(rtol=0, atol=1e-9)                                        # downstream of the SVD (reference: LAPACK + BLAS)
    if len(cons[f"{name}_core"]):
        got, _, _ = O.superpose_core(aln, coords, ref)
        np.testing.assert_allclose(np.concatenate(got), cons[f"{name}_sup_core"], **tol)
        got, _, _ = O.superpose_core(aln, coords, (ref + 1) % ch.n)
        np.testing.assert_allclose(np.concatenate(got), cons[f"{name}_sup_core_other"], **tol)
    if len(cons[f"{name}_sup_reference"]):
        got = O.superpose_reference(aln, coords, ref)[0]
        np.testing.assert_allclose(np.concatenate(got), cons[f"{name}_sup_reference"], **tol)
    else:
        with pytest.raises(AssertionError):
            O.superpose_reference(aln, coords, ref)
    if len(cons[f"{name}_sup_auto"]):
        np.testing.assert_allclose(np.concatenate(O.superpose(aln, coords)[2]), cons[f"{name}_sup_auto"], **tol)
    for mc in (50, 80):
        if len(cons[f"{name}_suprefs{mc}"]):
            np.testing.assert_allclose(np.concatenate(O.superpose_references(aln, coords, mc)), cons[f"{name}_suprefs{mc}"], **tol)


def test_consumers_text(cons):
    for key in ("special", "random"):
        M = cons[f"{key}_matrix"]
        names = [f"n{i}" for i in range(M.shape[0])] if key == "special" else [f"id{i}/chain{'A' * (i % 4)}" for i in range(40)]
        assert O.format_matrix(names, M) == cons[f"{key}_txt"].tobytes(), key
    for name in CC.CASES:
        names, seqs = [str(x) for x in cons[f"{name}_pnames"]], [str(x) for x in cons[f"{name}_seqs"]]
        assert O.format_fasta(names, seqs, cons[f"{name}_aln"]) == cons[f"{name}_fasta"].tobytes(), name
        assert O.format_matrix(names, cons[f"{name}_cg_distance"]) == cons[f"{name}_dist_txt"].tobytes(), name


def test_consumers_fast_mode_guide_matrix(cons):
    """make_count_matrix / braycurtis (multiple_alignment.py:128-145) against the reference's numba output: bit-exact, also on
    non-integer rows where the summation order matters."""
    off = np.concatenate([[0], np.cumsum(cons["bc_lengths"])])
    res = [cons["bc_indices"][off[p]:off[p + 1]] for p in range(len(cons["bc_lengths"]))]
    counts = O.count_matrix(res, 1024)
    assert np.array_equal(counts, cons["bc_counts"])
    assert np.array_equal(O.braycurtis(counts, counts), cons["bc_dist"])
    assert np.array_equal(O.braycurtis(cons["bc_x"], cons["bc_y"]), cons["bc_xy"])
    with pytest.raises(IndexError):
        O.count_matrix([np.array([0, 1024])], 1024)
