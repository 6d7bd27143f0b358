"""install() against the real reference module (imported with the I/O-only third-party modules mocked, like the oracle harness):
every attribute it replaces exists in the reference with a compatible signature.  Runs only where /root/reference is present
(the build container); the GPU box has no reference."""
import inspect

import pytest

from oracle import ref_harness

pytestmark = pytest.mark.skipif(not ref_harness.available(), reason="reference not present")


def test_install_replaces_existing_reference_attributes(monkeypatch):
    ma, dtw, sf, sup, helper, nj = ref_harness.load()
    from caretta_b200 import multiple_alignment as MA
    names = ["make_coverage_gap_distance_matrix", "get_reference_structures", "superpose", "superpose_core", "superpose_reference",
             "superpose_references", "make_rmsd_coverage_tm_matrix", "make_count_matrix", "braycurtis"]
    before = {n: getattr(ma, n) for n in names}
    methods = ["make_pairwise_matrix", "progressive_align", "to_sequence_alignment", "write_alignment"]
    before_m = {n: getattr(ma.MultipleAlignment, n) for n in methods}
    before_h, before_nj = helper.write_distance_matrix, nj.neighbor_joining
    try:
        for n in names:                                   # same parameter names as the reference's functions
            ref_fn = getattr(before[n], "py_func", before[n])
            ref_params = list(inspect.signature(ref_fn).parameters)
            got_params = list(inspect.signature(getattr(MA, n)).parameters)
            assert got_params[:len(ref_params)] == ref_params or ref_params[:len(got_params)] == got_params, (n, ref_params, got_params)
        for n in methods:
            ref_params = list(inspect.signature(before_m[n]).parameters)
            got_params = list(inspect.signature(getattr(MA.MultipleAlignment, n)).parameters)
            assert got_params == ref_params, (n, ref_params, got_params)
        assert list(inspect.signature(before_h).parameters) == list(inspect.signature(MA.write_distance_matrix).parameters)
        MA.install(ma)
        for n in names:
            assert getattr(ma, n) is getattr(MA, n), n
        for n in methods:
            assert getattr(ma.MultipleAlignment, n) is not before_m[n], n
        assert helper.write_distance_matrix is MA.write_distance_matrix
        assert nj.neighbor_joining is not before_nj
    finally:
        for n, f in before.items():
            setattr(ma, n, f)
        for n, f in before_m.items():
            setattr(ma.MultipleAlignment, n, f)
        helper.write_distance_matrix = before_h
        nj.neighbor_joining = before_nj


def test_mirror_signatures_of_the_sequence_interface():
    """Protein / SequenceBase methods, get_mean_weights, make_score_matrix: the reference's parameter names."""
    ma, dtw, sf, sup, helper, nj = ref_harness.load()
    from caretta_b200 import multiple_alignment as MA
    for cls, meths in ((ma.Protein, ["score_function", "mean_function"]), (ma.SequenceBase, ["score_function", "mean_function"])):
        for m in meths:
            ref_params = list(inspect.signature(getattr(cls, m)).parameters)
            got_params = list(inspect.signature(getattr(getattr(MA, cls.__name__), m)).parameters)
            assert got_params == ref_params, (cls.__name__, m, ref_params, got_params)
    assert [f.name for f in __import__("dataclasses").fields(MA.Protein)] == [f.name for f in __import__("dataclasses").fields(ma.Protein)]
    for fn, ref_fn in (("get_mean_weights", ma.get_mean_weights), ("make_score_matrix", sf.make_score_matrix.py_func),
                       ("get_gaussian_score", sf.get_gaussian_score.py_func), ("read_distance_matrix", helper.read_distance_matrix),
                       ("alignment_to_numpy", ma.alignment_to_numpy)):
        assert list(inspect.signature(getattr(MA, fn)).parameters) == list(inspect.signature(ref_fn).parameters), fn


def test_host_side_format_helpers_match_the_reference(tmp_path):
    """read_distance_matrix / alignment_to_numpy are file and string parsing on the host: same results as the reference's."""
    import numpy as np
    ma, dtw, sf, sup, helper, nj = ref_harness.load()
    from caretta_b200 import multiple_alignment as MA
    aln = {"a": "AC--GT-", "b/x": "-CCC--T", "c": "-------", "d": "ACDEFGH"}
    r, m = ma.alignment_to_numpy(aln), MA.alignment_to_numpy(aln)
    assert list(r) == list(m) and all(np.array_equal(r[k], m[k]) for k in r)
    names = ["p1/A", "q_2", "r3", "s/t/u"]
    M = np.random.default_rng(0).random((4, 4)) * 100
    fn = tmp_path / "m.txt"
    helper.write_distance_matrix(names, M, fn)
    a, b = helper.read_distance_matrix(fn), MA.read_distance_matrix(fn)
    assert a[0] == b[0] == ["p1", "q_2", "r3", "s"] and np.array_equal(a[1], b[1]) and b[1].dtype == np.float64
