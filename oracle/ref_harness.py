"""TEST INFRASTRUCTURE ONLY.  Imports the *unmodified* reference (TurtleTools/caretta at /root/reference)
in the build container so that golden vectors can be generated and the C restatement can be pinned.

/root/reference does not exist on the GPU box: nothing under tests/ -m gpu, smoke() or bench.py may
import this module.  Recipe: SURVEY.md Appendix C (I/O-only third-party modules are replaced by mocks;
no arithmetic lives in them on the hot path).
"""
import os
import sys
from unittest.mock import MagicMock

REFERENCE_ROOT = os.environ.get("CARETTA_REFERENCE", "/root/reference")


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
