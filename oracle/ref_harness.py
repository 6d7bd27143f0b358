"""TEST INFRASTRUCTURE ONLY.  Imports the *unmodified* reference (TurtleTools/caretta at /root/reference)
in the build container so that golden vectors can be generated and the C restatement can be pinned.

/root/reference does not exist on the GPU box; there the unmodified copy under oracle/_ref (build output of
oracle/build_ref.py, git-ignored) is imported instead -- only by bench.py's CPU legs and parity gate.  Recipe: SURVEY.md Appendix C (I/O-only third-party modules are replaced by mocks;
no arithmetic lives in them on the hot path).
"""
import os
import sys
from unittest.mock import MagicMock

from . import build_ref as _build_ref

# /root/reference in the build container; on the GPU box the unmodified copy build() left under oracle/_ref (build_ref.py)
REFERENCE_ROOT = _build_ref.root() or os.environ.get("CARETTA_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "caretta"))


def load():
    """Returns the reference modules (ma, dtw, score_functions, superposition_functions, helper, nj)."""
    if not available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    for m in ["Bio", "Bio.PDB", "Bio.PDB.ResidueDepth", "prody", "geometricus",
              "geometricus.protein_utility"]:
        sys.modules.setdefault(m, MagicMock(name=m))
    from caretta import multiple_alignment as ma, dynamic_time_warping as dtw, score_functions, \
        superposition_functions, helper, neighbor_joining as nj
    return ma, dtw, score_functions, superposition_functions, helper, nj


def proteins_from_chains(ma, chains):
    out = []
    for p in range(chains.n):
        t, c = chains.chain(p)
        out.append(ma.Protein(f"s{p}", t.copy(), c.copy(), "A" * len(t)))
    return out
