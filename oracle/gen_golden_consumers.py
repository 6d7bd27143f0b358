"""TEST INFRASTRUCTURE.  Golden vectors for the consumers of the multiple alignment, produced by
running the UNMODIFIED reference in the build container:  python oracle/gen_golden_consumers.py -> tests/golden/consumers.npz.

Inputs: the alignments of tests/golden/msa.npz (the reference's own multiple_align output on synthetic chains) and two
hand-made gappy alignments ("blocks": four groups of proteins on shifted column windows, so that get_reference_structures needs
more than one reference; "sparse": few gap-free columns, so that superpose() takes the reference branch).
Stored per case: make_coverage_gap_distance_matrix, get_reference_structures, superpose_core / superpose_reference / superpose /
superpose_references coordinates, the bytes of helper.write_distance_matrix and MultipleAlignment.write_alignment."""
import io
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_harness  # noqa: E402
from caretta_b200 import synth  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
LETTERS = np.array(list("ACDEFGHIKLMNPQRSTVWY"))


def hand_alignment(kind: str, seed: int):
    """(lengths, aln [N, A]) of a synthetic gappy alignment; residues of a protein appear in increasing order."""
    rng = np.random.default_rng(seed)
    if kind == "blocks":
        # four groups of three proteins on windows of 100 columns shifted by 40: neighbouring groups share 60 % of their
        # columns, second neighbours 20 %, so one reference cannot cover everybody at minimum_coverage = 50
        n, A = 12, 220
        rows = []
        for p in range(n):
            lo = 40 * (p // 3)
            keep = np.zeros(A, bool)
            keep[lo:lo + 100] = rng.random(100) < 0.93
            rows.append(keep)
    else:
        n, A = 7, 160
        rows = [rng.random(A) < 0.55 for _ in range(n)]
        for r in rows:
            r[:12] = True                                   # a few common columns so that every pair has > 3
    aln = -np.ones((n, A), np.int64)
    lengths = []
    for p, keep in enumerate(rows):
        k = int(keep.sum())
        aln[p, keep] = np.arange(k)
        lengths.append(k)
    return lengths, aln


def special_matrix():
    v = [0.0, -0.0, 1.0, 0.5, 0.00005, 0.00015, 0.00025, 0.12345, 0.12355, 2.5e-5, -2.5e-5, 1e-7, -1e-7, 0.99995, 9.99995, 123456.78905,
         1e15 + 0.5, 4503599627370496.5, 9007199254740993.0, 1e19, 9.3e18, 1e22, -1e22, 1.7976931348623157e308, 5e-324, float("nan"),
         -float("nan"), float("inf"), -float("inf"), 1234.56785, 0.1, 0.2, 0.30000000000000004, 65.4321, 7.00005, 3.99995, 2 ** 63, 2.0 ** 64,
         99999.99995, 0.49995, 1 / 3, 2 / 3, 1e4, 1e-4, 5e-5, 1.5e-4, 8.5e-4, 4.35, 4.45, 1.00005, 1.00015, 1.00025, 2.675]
    v = np.array(v + [0.0] * (56 - len(v)), dtype=np.float64)[:56]
    return v.reshape(7, 8)


def main():
    ma, dtw, sf, sup, helper, nj = ref_harness.load()
    msa = np.load(os.path.join(GOLD, "msa.npz"))
    out = {}
    cases = []
    for name in ("fam8", "ragged12"):
        ch = synth.make_chains(len(msa[f"{name}_lengths"]), list(msa[f"{name}_lengths"]), 10, seed=int(msa[f"{name}_seed"]),
                               family_size=int(msa[f"{name}_family"]))
        cases.append((name, [ch.chain(p)[1] for p in range(ch.n)], [ch.chain(p)[0] for p in range(ch.n)], msa[f"{name}_aln"], int(msa[f"{name}_seed"])))
    for kind, seed in (("blocks", 201), ("sparse", 202)):
        lengths, aln = hand_alignment(kind, seed)
        ch = synth.make_chains(len(lengths), lengths, 10, seed=seed, family_size=len(lengths))
        cases.append((kind, [ch.chain(p)[1] for p in range(ch.n)], [ch.chain(p)[0] for p in range(ch.n)], aln, seed))
        out[f"{kind}_lengths"] = np.array(lengths)
        out[f"{kind}_seed"] = seed
    out["names"] = np.array([c[0] for c in cases])
    for name, coords, tensors, aln, seed in cases:
        n = len(coords)
        rng = np.random.default_rng(seed + 1000)
        seqs = ["".join(LETTERS[rng.integers(0, 20, len(c))]) for c in coords]
        pnames = [f"prot_{name}_{p}{'x' * (p % 3)}" for p in range(n)]

        def proteins():
            return [ma.Protein(pnames[p], tensors[p].copy(), coords[p].copy(), seqs[p]) for p in range(n)]

        alignment = {pnames[p]: np.asarray(aln[p], dtype=np.int64) for p in range(n)}
        out[f"{name}_aln"] = np.asarray(aln, dtype=np.int64)
        out[f"{name}_seqs"] = np.array(seqs)
        out[f"{name}_pnames"] = np.array(pnames)
        dist, al = ma.make_coverage_gap_distance_matrix(np.asarray(aln, dtype=np.int64))
        out[f"{name}_cg_distance"], out[f"{name}_cg_aligning"] = dist, al
        for mc in (50, 80):
            first, refs, no_al = ma.get_reference_structures(alignment, mc)
            out[f"{name}_refs{mc}_first"] = np.array(pnames.index(first))
            out[f"{name}_refs{mc}_keys"] = np.array([pnames.index(k) for k in refs], dtype=np.int64)
            out[f"{name}_refs{mc}_off"] = np.cumsum([0] + [len(v) for v in refs.values()])
            out[f"{name}_refs{mc}_members"] = np.array([pnames.index(x) for v in refs.values() for x in v], dtype=np.int64)
            out[f"{name}_refs{mc}_noalign"] = np.array([pnames.index(x) for x in no_al], dtype=np.int64)
            try:
                P = ma.superpose_references(alignment, proteins(), mc)
                out[f"{name}_suprefs{mc}"] = np.concatenate([p.coordinates for p in P])
            except AssertionError:
                out[f"{name}_suprefs{mc}"] = np.zeros((0, 3))
        ref_name = sorted(alignment.keys(), key=lambda x: sum(1 for a in alignment[x] if a != -1), reverse=True)[0]
        out[f"{name}_reference"] = np.array(pnames.index(ref_name))
        core = np.array([i for i in range(aln.shape[1]) if -1 not in [alignment[k][i] for k in alignment]], dtype=np.int64)
        out[f"{name}_core"] = core
        if len(core):
            P = ma.superpose_core(alignment, proteins(), ref_name)
            out[f"{name}_sup_core"] = np.concatenate([p.coordinates for p in P])
            other = pnames[(pnames.index(ref_name) + 1) % n]
            P = ma.superpose_core(alignment, proteins(), other)
            out[f"{name}_sup_core_other"] = np.concatenate([p.coordinates for p in P])
        try:
            P = ma.superpose_reference(alignment, proteins(), ref_name)
            out[f"{name}_sup_reference"] = np.concatenate([p.coordinates for p in P])
        except AssertionError:
            out[f"{name}_sup_reference"] = np.zeros((0, 3))
        try:
            P = ma.superpose(alignment, proteins())
            out[f"{name}_sup_auto"] = np.concatenate([p.coordinates for p in P])
        except AssertionError:
            out[f"{name}_sup_auto"] = np.zeros((0, 3))
        with tempfile.TemporaryDirectory() as td:
            m = ma.MultipleAlignment(proteins())
            m.alignment = alignment
            m.write_alignment(os.path.join(td, "a.fasta"))
            out[f"{name}_fasta"] = np.frombuffer(open(os.path.join(td, "a.fasta"), "rb").read(), dtype=np.uint8)
            helper.write_distance_matrix(pnames, dist, os.path.join(td, "d.txt"))
            out[f"{name}_dist_txt"] = np.frombuffer(open(os.path.join(td, "d.txt"), "rb").read(), dtype=np.uint8)
        print(f"[gen-consumers] {name}: N={n} A={aln.shape[1]} core={len(core)} refs50={len(out[f'{name}_refs50_keys'])} "
              f"refs80={len(out[f'{name}_refs80_keys'])}")
    # text writer: special values and a random matrix with wide dynamic range
    with tempfile.TemporaryDirectory() as td:
        M = special_matrix()
        names = [f"n{i}" for i in range(M.shape[0])]
        helper.write_distance_matrix(names, M, os.path.join(td, "s.txt"))
        out["special_matrix"] = M
        out["special_txt"] = np.frombuffer(open(os.path.join(td, "s.txt"), "rb").read(), dtype=np.uint8)
        rng = np.random.default_rng(77)
        R = rng.normal(size=(40, 40)) * 10.0 ** rng.integers(-6, 7, size=(40, 40))
        R[rng.random((40, 40)) < 0.1] = 0.0
        T = np.round(rng.random((40, 40)) * 100, 4) + 0.00005                  # values near rounding ties
        R[:, 20:] = T[:, 20:]
        names = [f"id{i}/chain{'A' * (i % 4)}" for i in range(40)]
        helper.write_distance_matrix(names, R, os.path.join(td, "r.txt"))
        out["random_matrix"] = R
        out["random_txt"] = np.frombuffer(open(os.path.join(td, "r.txt"), "rb").read(), dtype=np.uint8)
    # fast-mode guide matrix: make_count_matrix + braycurtis (multiple_alignment.py:128-145) on synthetic shapemer indices
    # (alphabet 2^10 like model.output_dimension = 10), and braycurtis on non-integer rows (summation order matters there)
    import numba as nb
    rng = np.random.default_rng(91)
    K = 1024
    hot = rng.choice(K, 60, replace=False)
    res = nb.typed.List()
    flat = []
    for p in range(37):
        L = int(rng.integers(40, 300))
        r = np.where(rng.random(L) < 0.8, rng.choice(hot, L), rng.integers(0, K, L)).astype(np.int64)
        res.append(r); flat.append(r)
    counts = ma.make_count_matrix(res, K)
    out["bc_indices"] = np.concatenate(flat)
    out["bc_lengths"] = np.array([len(r) for r in flat])
    out["bc_counts"] = counts
    out["bc_dist"] = ma.braycurtis(counts, counts)
    X, Y = rng.random((23, 77)) * 3, rng.random((19, 77)) * 3
    out["bc_x"], out["bc_y"], out["bc_xy"] = X, Y, ma.braycurtis(X, Y)
    np.savez_compressed(os.path.join(GOLD, "consumers.npz"), **out)
    print("[gen-consumers] wrote", os.path.join(GOLD, "consumers.npz"), os.path.getsize(os.path.join(GOLD, "consumers.npz")), "bytes")


if __name__ == "__main__":
    main()
