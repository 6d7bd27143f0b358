"""Shared loader for the golden vectors of the alignment consumers (tests/golden/consumers.npz, made by
oracle/gen_golden_consumers.py from the unmodified reference)."""
import os

import numpy as np

from caretta_b200 import synth

G = os.path.join(os.path.dirname(__file__), "golden")
CASES = ("fam8", "ragged12", "blocks", "sparse")


def load():
    return np.load(os.path.join(G, "consumers.npz"))


def chains_of(name, gold):
    """The synthetic chains the golden case was generated from."""
    if name in ("fam8", "ragged12"):
        msa = np.load(os.path.join(G, "msa.npz"))
        lengths, seed, fam = list(msa[f"{name}_lengths"]), int(msa[f"{name}_seed"]), int(msa[f"{name}_family"])
    else:
        lengths, seed = list(gold[f"{name}_lengths"]), int(gold[f"{name}_seed"])
        fam = len(lengths)
    return synth.make_chains(len(lengths), lengths, 10, seed=seed, family_size=fam)


def reference_groups(gold, name, mc):
    keys, off, mem = gold[f"{name}_refs{mc}_keys"], gold[f"{name}_refs{mc}_off"], gold[f"{name}_refs{mc}_members"]
    return int(gold[f"{name}_refs{mc}_first"]), {int(k): [int(x) for x in mem[off[q]:off[q + 1]]] for q, k in enumerate(keys)}, \
        [int(x) for x in gold[f"{name}_refs{mc}_noalign"]]
