"""TEST INFRASTRUCTURE.  Golden vectors for neighbor joining, produced by running the UNMODIFIED
reference (caretta/neighbor_joining.py, numba) in the build container:  python oracle/gen_golden_nj.py
-> tests/golden/nj.npz.  Cases: random symmetric and asymmetric matrices of 3..60 nodes, an integer-valued matrix with
many exact ties (first row-major minimum), and the guide-tree input of BASELINE config 2: max(S) - S of the reference's
own 200 x 200 pairwise score matrix (multiple_alignment.py:501), rebuilt from tests/golden/c2_full.npz."""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_harness  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")


def main():
    nj = ref_harness.load()[5]
    rng = np.random.default_rng(1)
    out = {}
    names = []

    def add(name, D):
        t0 = time.time()
        tree, bl = nj.neighbor_joining(D.copy())
        out[f"{name}_D"] = D
        out[f"{name}_tree"] = tree
        out[f"{name}_bl"] = bl
        names.append(name)
        print(f"[gen-nj] {name}: n={D.shape[0]} rows={tree.shape[0]} {time.time() - t0:.1f}s")

    for n in (3, 4, 5, 8, 20, 60):
        for sym in (True, False):
            A = rng.random((n, n)) * 10
            if sym:
                A = (A + A.T) / 2
            np.fill_diagonal(A, 0)
            add(f"rand{n}_{'sym' if sym else 'asym'}", A)
    A = np.round(rng.random((30, 30)) * 6)
    A = (A + A.T) / 2
    np.fill_diagonal(A, 0)
    add("ties30", A)
    g = np.load(os.path.join(GOLD, "c2_full.npz"))
    n = 200
    S = np.zeros((n, n))
    ii, jj = np.triu_indices(n, 1)
    S[ii, jj] = g["score"]
    S[jj, ii] = g["score"]
    add("c2_guide", np.max(S) - S)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(GOLD, "nj.npz"), **out)
    print("[gen-nj] wrote nj.npz")


if __name__ == "__main__":
    main()
