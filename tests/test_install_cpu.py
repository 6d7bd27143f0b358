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
