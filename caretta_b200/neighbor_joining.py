"""Mirror of caretta/neighbor_joining.py (TurtleTools/caretta 0.2.0) backed by the CUDA engine: same function name,
argument and return types, bit-identical output (tests/test_gpu_nj.py).

    reference                                               here
    ------------------------------------------------------  ---------------------------------------------------
    neighbor_joining(distance_matrix) -> (tree uint64       neighbor_joining(distance_matrix): crt_neighbor_joining through
      [k,2], branch_lengths float64 [k,1]), :17-99           the C ABI; O(N^3) instead of the reference's O(N^4)

Called by the reference at multiple_alignment.py:277 on ``max(S) - S`` of the pairwise score matrix.
"""
from __future__ import annotations

import numpy as np

from . import multiple_alignment as _ma


def neighbor_joining(distance_matrix: np.ndarray):
    return _ma.get_engine().neighbor_joining(distance_matrix)


def install(reference_neighbor_joining_module) -> None:
    """Replaces the reference module's neighbor_joining in place (callers that imported the module, like
    multiple_alignment.py:9 `from caretta import neighbor_joining as nj`, pick it up)."""
    reference_neighbor_joining_module.neighbor_joining = neighbor_joining
