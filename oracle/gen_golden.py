"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, numba) in the build container on seeded synthetic inputs.

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden.py [--skip-c2]

The fixtures are committed; this script is committed with them so they can be regenerated and audited.
Inputs are re-created from seeds by caretta_b200.synth (the same generator the tests use), so only outputs are
stored (plus a checksum of the inputs to catch generator drift).
Reference call sites exercised: multiple_alignment.py:158-170 (make_pairwise_matrix), :321-349
(Protein.score_function), :255-285 (multiple_align), :1000-1055 (make_rmsd_coverage_tm_matrix),
dynamic_time_warping.py (all), score_functions.py (all), superposition_functions.py (all), helper.py:12-53.
"""
import argparse
import hashlib
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from caretta_b200 import synth  # noqa: E402
from oracle import ref_harness  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
GT, GC = 7.0, 0.03
PARAMS = dict(flexible=False, gamma_tensor=GT, gamma_coords=GC, verbose=False)


def digest(*arrays) -> str:
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def ragged(list_of_arrays, dtype):
    off = np.zeros(len(list_of_arrays) + 1, np.int64)
    off[1:] = np.cumsum([len(a) for a in list_of_arrays])
    flat = np.concatenate([np.asarray(a, dtype) for a in list_of_arrays]) if list_of_arrays else np.zeros(0, dtype)
    return flat, off


def ref_pair(mods, t1, c1, t2, c2):
    """The pair recipe, composed from reference functions exactly as Protein.score_function does
    (multiple_alignment.py:328-349) but keeping the intermediates."""
    ma, dtw, sf, sup, helper, nj = mods
    S_T = sf.make_score_matrix(t1, t2, sf.get_gaussian_score, GT)
    a1, a2, sc1 = dtw.smith_waterman(np.arange(S_T.shape[0]), np.arange(S_T.shape[1]), S_T, 0.)
    p1, p2 = helper.get_common_positions(a1, a2)
    R = np.eye(3)
    if len(p1) <= 3:
        w1, w2 = np.array(c1), np.array(c2)
    else:
        w1, w2, _ = sup.paired_svd_superpose_with_subset(c1, c2, c1[p1], c2[p2])
        R, _ = sup.paired_svd_superpose(c1[p1], c2[p2])
    S_C = sf.make_score_matrix(w1, w2, sf.get_gaussian_score, GC)
    score = dtw.smith_waterman_score(np.arange(S_C.shape[0]), np.arange(S_C.shape[1]), S_C)
    rmsd = tm = 0.0
    if len(p1) >= 1:
        k1, k2 = c1[p1], c2[p2]
        if len(p1) > 3:
            rot, tran = sup.paired_svd_superpose(k1, k2)
            k2 = sup.apply_rotran(k2, rot, tran)
        rmsd = sf.get_rmsd(k1, k2)
        tm = ma.tm_score(k1, k2, len(t1), len(t2))
    return dict(S_T=S_T, a1=a1, a2=a2, sc1=sc1, p1=p1, p2=p2, R=R, S_C=S_C, score=score, rmsd=rmsd, tm=tm)


_MODS = None
_CH = None


def _c2_worker(args):
    lo, hi, pairs = args
    out = []
    for q in range(lo, hi):
        i, j = pairs[q]
        ti, ci = _CH.chain(i)
        tj, cj = _CH.chain(j)
        r = ref_pair(_MODS, ti, ci, tj, cj)
        out.append((q, r["a1"].astype(np.int16), r["a2"].astype(np.int16), r["score"], r["sc1"], len(r["p1"]),
                    r["rmsd"], r["tm"]))
    return out


def main():
    global _MODS, _CH
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-c2", action="store_true")
    args = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    t0 = time.time()
    mods = ref_harness.load()
    ma, dtw, sf, sup, helper, nj = mods
    _MODS = mods
    ma.trigger_numba_compilation()
    print(f"[gen] reference imported + JIT warm-up {time.time() - t0:.1f}s")

    # ---------------------------------------------------------------- KATs (SURVEY.md Appendix B and more)
    kat = {}
    S = np.array([[1, .1, .1], [.1, 1, .1]], dtype=np.float64)
    a1, a2, sc = dtw.dtw_align(np.arange(2), np.arange(3), S, 1.0, 0.01)
    M, B = dtw._make_dtw_matrix(np.arange(2), np.arange(3), S, 1.0, 0.01)
    kat.update(dtw0_S=S, dtw0_a1=a1, dtw0_a2=a2, dtw0_sc=sc, dtw0_M=M, dtw0_B=B)
    for name, S, go, ge in [("dtw1", np.ones((3, 3)), 0., 0.), ("dtw2", np.full((3, 4), .5), 1., .01),
                            ("dtw3", np.full((4, 2), .25), 1., .01), ("dtw4", np.zeros((3, 3)), 1., .01),
                            ("dtw5", np.full((1, 1), .7), 1., .01), ("dtw6", np.full((1, 5), .7), 1., .01)]:
        a1, a2, sc = dtw.dtw_align(np.arange(S.shape[0]), np.arange(S.shape[1]), S, go, ge)
        kat.update({f"{name}_S": S, f"{name}_a1": a1, f"{name}_a2": a2, f"{name}_sc": sc,
                    f"{name}_go": go, f"{name}_ge": ge})
    for name, S in [("sw0", np.ones((3, 3))), ("sw1", np.full((3, 4), .5)), ("sw2", np.full((4, 3), .5)),
                    ("sw3", np.array([[0., 0., 0.], [0., .5, 0.], [0., 0., .25]])),
                    ("sw4", np.array([[0., 0., .3], [0., 0., 0.], [.2, 0., 0.]]))]:
        a1, a2, sc = dtw.smith_waterman(np.arange(S.shape[0]), np.arange(S.shape[1]), S, 0.)
        kat.update({f"{name}_S": S, f"{name}_a1": a1, f"{name}_a2": a2, f"{name}_sc": sc,
                    f"{name}_score_only": dtw.smith_waterman_score(np.arange(S.shape[0]), np.arange(S.shape[1]), S)})
    try:
        dtw.smith_waterman(np.arange(3), np.arange(3), np.zeros((3, 3)), 0.)
        kat["sw_zero_raises"] = False
    except Exception:
        kat["sw_zero_raises"] = True
    kat["sw_zero_score_only"] = dtw.smith_waterman_score(np.arange(3), np.arange(3), np.zeros((3, 3)))
    kat["tm_zero"] = ma.tm_score(np.zeros((2, 3)), np.zeros((2, 3)), 2, 2)
    rng = np.random.default_rng(77)
    x, y = rng.normal(size=(17, 3)) * 5, rng.normal(size=(17, 3)) * 5
    kat.update(tm_x=x, tm_y=y, tm_val=ma.tm_score(x, y, 40, 23), rmsd_val=sf.get_rmsd(x, y))
    R, t = sup.paired_svd_superpose(x, y)
    kat.update(kab_R=R, kab_t=t, kab_applied=sup.apply_rotran(y, R, t))
    ym = y.copy(); ym[:, 0] = -ym[:, 0]   # mirror image -> reflection branch
    Rm, tm_ = sup.paired_svd_superpose(x, ym)
    kat.update(kab_mirror_y=ym, kab_mirror_R=Rm, kab_mirror_t=tm_)
    aa = np.array([0, -1, 1, 2, -1, 3]); bb = np.array([0, 1, -1, 2, 3, -1])
    p1, p2 = helper.get_common_positions(aa, bb)
    kat.update(cp_a=aa, cp_b=bb, cp_p1=p1, cp_p2=p2)
    np.savez_compressed(os.path.join(GOLD, "kat.npz"), **kat)
    print(f"[gen] kat.npz ({len(kat)} arrays)")

    # ---------------------------------------------------------------- small mixed-length set, full intermediates
    lengths = [30, 47, 35, 52, 3, 4, 61, 40, 5, 44]
    ch = synth.make_chains(len(lengths), lengths, 10, seed=11, family_size=4)
    P = ref_harness.proteins_from_chains(ma, ch)
    Sfull = ma.MultipleAlignment(P).make_pairwise_matrix(dict(PARAMS))
    out = dict(lengths=np.array(lengths), seed=11, family_size=4, d=10, input_digest=digest(ch.coords, ch.tensors),
               score_matrix=Sfull)
    a1s, a2s, sc, sc1, nc, Rs, rm, tmv, pi, pj = [], [], [], [], [], [], [], [], [], []
    keep = {}
    for i in range(ch.n - 1):
        for j in range(i + 1, ch.n):
            ti, ci = ch.chain(i); tj, cj = ch.chain(j)
            r = ref_pair(mods, ti, ci, tj, cj)
            assert r["score"] == Sfull[i, j], (i, j)   # our composition == the reference driver, bit for bit
            pi.append(i); pj.append(j)
            a1s.append(r["a1"]); a2s.append(r["a2"]); sc.append(r["score"]); sc1.append(r["sc1"])
            nc.append(len(r["p1"])); Rs.append(r["R"]); rm.append(r["rmsd"]); tmv.append(r["tm"])
            if (i, j) in [(0, 1), (2, 3), (6, 9), (4, 5), (1, 8)]:
                keep[f"ST_{i}_{j}"] = r["S_T"]; keep[f"SC_{i}_{j}"] = r["S_C"]
    a1f, aoff = ragged(a1s, np.int16)
    a2f, _ = ragged(a2s, np.int16)
    out.update(pi=np.array(pi, np.int32), pj=np.array(pj, np.int32), aln1=a1f, aln2=a2f, aln_off=aoff,
               score=np.array(sc), score1=np.array(sc1), ncommon=np.array(nc, np.int32), R=np.array(Rs),
               rmsd=np.array(rm), tm=np.array(tmv), **keep)
    # full downstream chain on the chains long enough for the MSA (>= 3 common positions asserted by the reference)
    sel = [0, 1, 2, 3, 6, 7, 9]
    Psel = [P[q] for q in sel]
    msa = ma.MultipleAlignment(Psel)
    Ssel = msa.make_pairwise_matrix(dict(PARAMS))
    D = Ssel.max() - Ssel
    msa.multiple_align(D, gap_open_penalty=1.0, gap_extend_penalty=0.01, consensus_weight=1.0, gamma_weight=1.0,
                       score_function_params=dict(PARAMS), mean_function_params={})
    aln = np.array([msa.alignment[p.name] for p in Psel], dtype=np.int64)
    r_, c_, t_ = ma.make_rmsd_coverage_tm_matrix(msa.alignment, Psel, superpose_first=False)
    out.update(msa_sel=np.array(sel), msa_aln=aln, msa_rmsd=r_, msa_cov=c_, msa_tm=t_, msa_tree=msa.tree,
               msa_branch_lengths=msa.branch_lengths, msa_D=D)
    np.savez_compressed(os.path.join(GOLD, "pairs_small.npz"), **out)
    print(f"[gen] pairs_small.npz  pairs={len(pi)}  {time.time() - t0:.1f}s")

    # ---------------------------------------------------------------- affine DTW on realistic score matrices
    dt = {}
    rng = np.random.default_rng(5)
    cases = [(0, 1), (2, 3), (6, 9), (3, 6), (7, 9), (4, 5), (5, 8), (0, 4)]
    for q, (i, j) in enumerate(cases):
        ti, ci = ch.chain(i); tj, cj = ch.chain(j)
        S = P[i].score_function(P[j], **PARAMS)
        # consensus-weight term exactly as progressive_align adds it (multiple_alignment.py:207-210), weights = 1
        S = S + sf.make_score_matrix(np.ones((len(ti), 1)), np.ones((len(tj), 1)), sf.get_gaussian_score, 1.0)
        for go, ge in [(1.0, 0.01), (0.5, 0.1), (0.0, 0.0)]:
            a1, a2, sc_ = dtw.dtw_align(np.arange(S.shape[0]), np.arange(S.shape[1]), S, go, ge)
            key = f"c{q}_{go}_{ge}"
            dt[f"{key}_a1"] = a1.astype(np.int16); dt[f"{key}_a2"] = a2.astype(np.int16); dt[f"{key}_sc"] = sc_
        dt[f"c{q}_S"] = S
    for q in range(6):   # random matrices incl. ragged / tiny shapes and exact ties (quantised scores)
        n, m = [(1, 1), (1, 7), (9, 2), (13, 17), (25, 25), (31, 8)][q]
        S = np.round(rng.random((n, m)) * 4) / 4
        a1, a2, sc_ = dtw.dtw_align(np.arange(n), np.arange(m), S, 1.0, 0.01)
        dt[f"r{q}_S"] = S; dt[f"r{q}_a1"] = a1.astype(np.int16); dt[f"r{q}_a2"] = a2.astype(np.int16); dt[f"r{q}_sc"] = sc_
        a1, a2, sc_ = dtw.smith_waterman(np.arange(n), np.arange(m), S + 0.0, 0.)
        dt[f"r{q}_sw_a1"] = a1.astype(np.int16); dt[f"r{q}_sw_a2"] = a2.astype(np.int16); dt[f"r{q}_sw_sc"] = sc_
    dt["n_cases"] = len(cases)
    np.savez_compressed(os.path.join(GOLD, "dtw.npz"), **dt)
    print(f"[gen] dtw.npz  {time.time() - t0:.1f}s")

    # ---------------------------------------------------------------- C1: CA coords of the reference's sample PDBs
    td = os.path.join(ref_harness.REFERENCE_ROOT, "test_data")
    names = ["1kdu", "1pk4", "1pkr"]
    cas = [synth.read_ca_coords(os.path.join(td, f"{n}.pdb")) for n in names]
    rng = np.random.default_rng(1)
    lmax = max(len(c) for c in cas)
    base = rng.normal(0.0, 0.3, size=(lmax, 10))   # surrogate tensors (geometricus is not installable here)
    tens = [base[:len(c)] + rng.normal(0.0, 0.02, size=(len(c), 10)) for c in cas]
    Pc1 = [ma.Protein(n, t, c, "A" * len(c)) for n, t, c in zip(names, tens, cas)]
    Sc1 = ma.MultipleAlignment(Pc1).make_pairwise_matrix(dict(PARAMS))
    c1 = dict(names=np.array(names), score_matrix=Sc1)
    for n, t, c in zip(names, tens, cas):
        c1[f"ca_{n}"] = c; c1[f"tensors_{n}"] = t
    np.savez_compressed(os.path.join(GOLD, "c1_test_data.npz"), **c1)
    print(f"[gen] c1_test_data.npz lengths={[len(c) for c in cas]}  {time.time() - t0:.1f}s")

    # ---------------------------------------------------------------- C2 in full: 200 x 80, 19 900 pairs
    if not args.skip_c2:
        _CH = synth.config("C2")
        pairs = [(i, j) for i in range(_CH.n - 1) for j in range(i + 1, _CH.n)]
        nw = max(1, (os.cpu_count() or 2))
        step = (len(pairs) + nw * 8 - 1) // (nw * 8)
        jobs = [(lo, min(lo + step, len(pairs)), pairs) for lo in range(0, len(pairs), step)]
        ctx = mp.get_context("fork")   # fork after JIT warm-up
        ref_pair(mods, *_CH.chain(0), *_CH.chain(1))
        with ctx.Pool(nw) as pool:
            res = [r for chunk in pool.map(_c2_worker, jobs) for r in chunk]
        res.sort(key=lambda r: r[0])
        a1f, aoff = ragged([r[1] for r in res], np.int16)
        a2f, _ = ragged([r[2] for r in res], np.int16)
        np.savez_compressed(os.path.join(GOLD, "c2_full.npz"), input_digest=digest(_CH.coords, _CH.tensors),
                            aln1=a1f.astype(np.int8), aln2=a2f.astype(np.int8), aln_off=aoff.astype(np.int32),
                            score=np.array([r[3] for r in res]), score1=np.array([r[4] for r in res]),
                            ncommon=np.array([r[5] for r in res], np.int16),
                            rmsd=np.array([r[6] for r in res]), tm=np.array([r[7] for r in res]))
        print(f"[gen] c2_full.npz pairs={len(res)}  {time.time() - t0:.1f}s")


if __name__ == "__main__":
    main()
