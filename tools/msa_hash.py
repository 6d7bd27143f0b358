"""Hash of the progressive alignment of N synthetic chains x L (bit-level comparison between build / environment variants).
python tools/msa_hash.py [N] [L]"""
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from caretta_b200 import multiple_alignment as MA, neighbor_joining as NJ, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
L = int(sys.argv[2]) if len(sys.argv) > 2 else 300
ch = synth.make_chains(n, L, 10, seed=3, family_size=20)
msa = MA.StructureMultiple.from_chains(ch)
prm = dict(MA.DEFAULT_SCORE_PARAMS)
S = msa.make_pairwise_matrix(prm)
tree, _ = NJ.neighbor_joining(S.max() - S)
aln = msa.progressive_align(tree, 1.0, 0.01, 1.0, 0.03, prm, dict(flexible=False))
h = hashlib.sha256()
for k in sorted(aln):
    h.update(k.encode())
    h.update(np.ascontiguousarray(aln[k], dtype=np.int64).tobytes())
last = msa.final_sequences[-1]
h.update(np.ascontiguousarray(last.tensors).tobytes())
h.update(np.ascontiguousarray(last.coordinates).tobytes())
print(f"N={n} L={L} alignment length {len(next(iter(aln.values())))} sha256 {h.hexdigest()[:24]}")
