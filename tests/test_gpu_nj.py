"""Neighbor joining on the device against the golden vectors produced by the unmodified
reference (oracle/gen_golden_nj.py) and against the pinned oracle: tree rows and branch lengths bit-identical."""
import os

import numpy as np
import pytest

from caretta_b200 import engine, neighbor_joining as NJ
from oracle import oracle as O

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module", params=["default", "inplace", "two_buffer"])
def eng(request):
    """default: the in-place kernels from 512 nodes on; inplace / two_buffer: one path for every size (the golden cases are small)."""
    old = os.environ.get("CARETTA_B200_NJ_INPLACE_MIN")
    if request.param != "default":
        os.environ["CARETTA_B200_NJ_INPLACE_MIN"] = "4" if request.param == "inplace" else "0"
    e = engine.Engine()
    yield e
    e.close()
    if old is None:
        os.environ.pop("CARETTA_B200_NJ_INPLACE_MIN", None)
    else:
        os.environ["CARETTA_B200_NJ_INPLACE_MIN"] = old


def test_nj_golden_bit_exact(eng):
    g = np.load(os.path.join(G, "nj.npz"))
    for name in [str(n) for n in g["names"]]:
        tree, bl = eng.neighbor_joining(g[f"{name}_D"])
        assert tree.dtype == np.uint64 and bl.dtype == np.float64 and bl.shape == (tree.shape[0], 1)
        assert np.array_equal(tree, g[f"{name}_tree"]), name
        assert np.array_equal(bl, g[f"{name}_bl"]), name


def test_nj_larger_vs_oracle(eng):
    rng = np.random.default_rng(11)
    for n in (257, 600, 1100):
        A = rng.random((n, n)) * 3
        A = (A + A.T) / 2
        np.fill_diagonal(A, 0)
        tree, bl = eng.neighbor_joining(A)
        to, bo = O.neighbor_joining(A)
        assert np.array_equal(tree, to) and np.array_equal(bl, bo), n


def test_nj_mirror_and_errors(eng):
    A = np.array([[0., 5, 9, 9, 8], [5, 0, 10, 10, 9], [9, 10, 0, 8, 7], [9, 10, 8, 0, 3], [8, 9, 7, 3, 0]])
    tree, bl = NJ.neighbor_joining(A)
    to, bo = O.neighbor_joining(A)
    assert np.array_equal(tree, to) and np.array_equal(bl, bo)
    with pytest.raises(IndexError):
        eng.neighbor_joining(np.zeros((2, 2)))
    with pytest.raises(ValueError):
        eng.neighbor_joining(np.zeros((3, 4)))
