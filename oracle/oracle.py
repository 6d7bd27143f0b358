"""TEST INFRASTRUCTURE ONLY -- ctypes front end of oracle/libcaretta_oracle.so (the CPU restatement of the
reference's pair path, see caretta_oracle.c for the reference file:line map).

Importers allowed: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference legs.
The product package (caretta_b200/) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libcaretta_oracle.so")
_lib = None

_D = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_I64 = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_I32 = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_P = C.c_void_p


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "caretta_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.crt_o_rbf_matrix.argtypes = [_D, _D, C.c_int, C.c_int, C.c_int, C.c_double, _D]
        L.crt_o_rbf_matrix.restype = None
        L.crt_o_smith_waterman_score.argtypes = [_D, C.c_int, C.c_int, C.c_double]
        L.crt_o_smith_waterman_score.restype = C.c_double
        L.crt_o_smith_waterman.argtypes = [_D, C.c_int, C.c_int, C.c_double, _I64, _I64,
                                           C.POINTER(C.c_int64), C.POINTER(C.c_double)]
        L.crt_o_smith_waterman.restype = C.c_int
        L.crt_o_dtw_align.argtypes = [_D, C.c_int, C.c_int, C.c_double, C.c_double, _I64, _I64,
                                      C.POINTER(C.c_int64), C.POINTER(C.c_double), _P, _P]
        L.crt_o_dtw_align.restype = None
        L.crt_o_common_positions.argtypes = [_I64, _I64, C.c_int64, _I64, _I64]
        L.crt_o_common_positions.restype = C.c_int64
        L.crt_o_kabsch.argtypes = [_D, _D, C.c_int64, _D, _D]
        L.crt_o_kabsch.restype = None
        L.crt_o_apply_rotran.argtypes = [_D, C.c_int64, _D, _D, _D]
        L.crt_o_apply_rotran.restype = None
        L.crt_o_superpose_with_subset.argtypes = [_D, C.c_int64, _D, C.c_int64, _D, _D, C.c_int64, _D, _D, _P]
        L.crt_o_superpose_with_subset.restype = None
        L.crt_o_rmsd.argtypes = [_D, _D, C.c_int64]
        L.crt_o_rmsd.restype = C.c_double
        L.crt_o_tm_score.argtypes = [_D, _D, C.c_int64, C.c_int64, C.c_int64]
        L.crt_o_tm_score.restype = C.c_double
        L.crt_o_pair.argtypes = [_D, _D, C.c_int, _D, _D, C.c_int, C.c_int, C.c_double, C.c_double,
                                 C.POINTER(C.c_double), _P, _P, _P, _P, _P, _P, _P, _P]
        L.crt_o_pair.restype = C.c_int
        L.crt_o_pair_f32model.argtypes = [_D, _D, C.c_int, _D, _D, C.c_int, C.c_int, C.c_double, C.c_double,
                                          C.c_int, C.POINTER(C.c_double), _I64, _I64, C.POINTER(C.c_int64)]
        L.crt_o_pair_f32model.restype = C.c_int
        L.crt_o_pairwise_list.argtypes = [_D, _D, _I64, C.c_int, _I32, _I32, C.c_int64, C.c_double, C.c_double,
                                          C.c_int, _D, _P, _P, _P, _P]
        L.crt_o_pairwise_list.restype = None
        L.crt_o_pairwise_all.argtypes = [_D, _D, _I64, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, _D]
        L.crt_o_pairwise_all.restype = None
        L.crt_o_rmsd_cov_tm.argtypes = [_I64, C.c_int, C.c_int64, _D, _I64, _D, _D, _D]
        L.crt_o_rmsd_cov_tm.restype = C.c_int
        L.crt_o_num_threads.restype = C.c_int
        L.crt_o_score_matrix.argtypes = [_D, _D, C.c_int, _D, _D, C.c_int, C.c_int, C.c_double, C.c_double, _D]
        L.crt_o_score_matrix.restype = C.c_int
        L.crt_o_neighbor_joining.argtypes = [_D, C.c_int, np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS"), _D]
        L.crt_o_neighbor_joining.restype = C.c_int64
        L.crt_o_coverage_gap_matrix.argtypes = [_I64, C.c_int, C.c_int64, _D, _I32]
        L.crt_o_coverage_gap_matrix.restype = C.c_int
        L.crt_o_format_matrix.argtypes = [_D, C.c_int, C.c_int, C.c_char_p, _I64, _P]
        L.crt_o_format_matrix.restype = C.c_int64
        L.crt_o_count_matrix.argtypes = [_I64, _I64, C.c_int, C.c_int, _D]
        L.crt_o_count_matrix.restype = C.c_int
        L.crt_o_braycurtis.argtypes = [_D, C.c_int, _D, C.c_int, C.c_int, _D]
        L.crt_o_braycurtis.restype = None
        _lib = L
    return _lib


def _c(a, dt=np.float64):
    return np.ascontiguousarray(a, dtype=dt)


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def rbf_matrix(x, y, gamma: float) -> np.ndarray:
    x, y = _c(x), _c(y)
    S = np.empty((x.shape[0], y.shape[0]))
    lib().crt_o_rbf_matrix(x, y, x.shape[0], y.shape[0], x.shape[1], gamma, S)
    return S


def smith_waterman_score(S, gap: float = 0.0) -> float:
    S = _c(S)
    return float(lib().crt_o_smith_waterman_score(S, S.shape[0], S.shape[1], gap))


def smith_waterman(S, gap: float = 0.0):
    """Returns (aln1, aln2, score); raises ValueError where the reference raises (no cell > 0)."""
    S = _c(S)
    n, m = S.shape
    a1 = np.empty(n + m + 1, np.int64)
    a2 = np.empty(n + m + 1, np.int64)
    ln, sc = C.c_int64(0), C.c_double(0)
    rc = lib().crt_o_smith_waterman(S, n, m, gap, a1, a2, C.byref(ln), C.byref(sc))
    if rc != 0:
        raise ValueError("smith_waterman: no positive cell (reference raises here)")
    return a1[:ln.value].copy(), a2[:ln.value].copy(), sc.value


def dtw_align(S, gap_open: float, gap_extend: float, want_matrices: bool = False):
    S = _c(S)
    n, m = S.shape
    a1 = np.empty(n + m + 1, np.int64)
    a2 = np.empty(n + m + 1, np.int64)
    ln, sc = C.c_int64(0), C.c_double(0)
    M = np.empty((n + 1, m + 1, 3)) if want_matrices else None
    B = np.empty((n + 1, m + 1, 3), np.int64) if want_matrices else None
    lib().crt_o_dtw_align(S, n, m, gap_open, gap_extend, a1, a2, C.byref(ln), C.byref(sc), _ptr(M), _ptr(B))
    if want_matrices:
        return a1[:ln.value].copy(), a2[:ln.value].copy(), sc.value, M, B
    return a1[:ln.value].copy(), a2[:ln.value].copy(), sc.value


def common_positions(a1, a2):
    a1, a2 = _c(a1, np.int64), _c(a2, np.int64)
    p1 = np.empty(max(len(a1), 1), np.int64)
    p2 = np.empty(max(len(a1), 1), np.int64)
    c = lib().crt_o_common_positions(a1, a2, len(a1), p1, p2)
    return p1[:c].copy(), p2[:c].copy()


def kabsch(x1, x2) -> Tuple[np.ndarray, np.ndarray]:
    x1, x2 = _c(x1), _c(x2)
    R, t = np.empty((3, 3)), np.empty(3)
    lib().crt_o_kabsch(x1, x2, x1.shape[0], R, t)
    return R, t


def apply_rotran(x, R, t):
    x = _c(x)
    out = np.empty_like(x)
    lib().crt_o_apply_rotran(x, x.shape[0], _c(R), _c(t), out)
    return out


def superpose_with_subset(c1, c2, k1, k2):
    c1, c2, k1, k2 = _c(c1), _c(c2), _c(k1), _c(k2)
    o1, o2, R = np.empty_like(c1), np.empty_like(c2), np.empty((3, 3))
    lib().crt_o_superpose_with_subset(c1, c1.shape[0], c2, c2.shape[0], k1, k2, k1.shape[0], o1, o2, _ptr(R))
    return o1, o2, R


def rmsd(x, y) -> float:
    x, y = _c(x), _c(y)
    return float(lib().crt_o_rmsd(x, y, x.shape[0]))


def tm_score(x, y, l1: int, l2: int) -> float:
    x, y = _c(x), _c(y)
    return float(lib().crt_o_tm_score(x, y, x.shape[0], l1, l2))


def pair(t1, c1, t2, c2, gamma_t: float = 7.0, gamma_c: float = 0.03) -> dict:
    """Full pair recipe; returns a dict with score, aln1, aln2, ncommon, R, rmsd, tm, score1, status."""
    t1, c1, t2, c2 = _c(t1), _c(c1), _c(t2), _c(c2)
    n, m, d = t1.shape[0], t2.shape[0], t1.shape[1]
    a1 = np.empty(n + m + 1, np.int64)
    a2 = np.empty(n + m + 1, np.int64)
    ln, nc = np.zeros(1, np.int64), np.zeros(1, np.int64)
    R = np.empty((3, 3))
    sc = C.c_double(0)
    rr, tt, s1 = np.zeros(1), np.zeros(1), np.zeros(1)
    st = lib().crt_o_pair(t1, c1, n, t2, c2, m, d, gamma_t, gamma_c, C.byref(sc), _ptr(a1), _ptr(a2), _ptr(ln),
                          _ptr(nc), _ptr(R), _ptr(rr), _ptr(tt), _ptr(s1))
    return dict(score=sc.value, aln1=a1[:ln[0]].copy(), aln2=a2[:ln[0]].copy(), ncommon=int(nc[0]), R=R,
                rmsd=float(rr[0]), tm=float(tt[0]), score1=float(s1[0]), status=int(st))


def pair_f32model(t1, c1, t2, c2, gamma_t: float = 7.0, gamma_c: float = 0.03, variant: int = 1):
    t1, c1, t2, c2 = _c(t1), _c(c1), _c(t2), _c(c2)
    n, m, d = t1.shape[0], t2.shape[0], t1.shape[1]
    a1 = np.empty(n + m + 1, np.int64)
    a2 = np.empty(n + m + 1, np.int64)
    ln, sc = C.c_int64(0), C.c_double(0)
    lib().crt_o_pair_f32model(t1, c1, n, t2, c2, m, d, gamma_t, gamma_c, variant, C.byref(sc), a1, a2, C.byref(ln))
    return a1[:ln.value].copy(), a2[:ln.value].copy(), sc.value


def pairwise_list(coords, tensors, offsets, pi, pj, gamma_t=7.0, gamma_c=0.03, nthreads=0, extras=True):
    coords, tensors = _c(coords), _c(tensors)
    offsets = _c(offsets, np.int64)
    pi, pj = _c(pi, np.int32), _c(pj, np.int32)
    n = len(pi)
    score = np.empty(n)
    rm = np.empty(n) if extras else None
    tm = np.empty(n) if extras else None
    nc = np.empty(n, np.int32) if extras else None
    st = np.empty(n, np.int32) if extras else None
    lib().crt_o_pairwise_list(coords, tensors, offsets, tensors.shape[1], pi, pj, n, gamma_t, gamma_c, nthreads,
                              score, _ptr(rm), _ptr(tm), _ptr(nc), _ptr(st))
    return dict(score=score, rmsd=rm, tm=tm, ncommon=nc, status=st)


def pairwise_all(coords, tensors, offsets, gamma_t=7.0, gamma_c=0.03, nthreads=0) -> np.ndarray:
    coords, tensors = _c(coords), _c(tensors)
    offsets = _c(offsets, np.int64)
    N = len(offsets) - 1
    out = np.empty((N, N))
    lib().crt_o_pairwise_all(coords, tensors, offsets, N, tensors.shape[1], gamma_t, gamma_c, nthreads, out)
    return out


def pairwise_all_flexible(tensors, offsets, gamma_t=7.0) -> np.ndarray:
    """make_pairwise_matrix (multiple_alignment.py:158-170) with score_function_params flexible=True (:323-326):
    smith_waterman_score of the tensor Gaussian of every pair (small cases: a Python loop over the pairs)."""
    tensors, offsets = _c(tensors), _c(offsets, np.int64)
    N = len(offsets) - 1
    out = np.zeros((N, N))
    for i in range(N - 1):
        for j in range(i + 1, N):
            out[i, j] = out[j, i] = smith_waterman_score(
                rbf_matrix(tensors[offsets[i]:offsets[i + 1]], tensors[offsets[j]:offsets[j + 1]], gamma_t), 0.0)
    return out


def rmsd_cov_tm(aln, coords, offsets):
    aln = _c(aln, np.int64)
    coords = _c(coords)
    offsets = _c(offsets, np.int64)
    N, A = aln.shape
    r, c, t = np.empty((N, N)), np.empty((N, N)), np.empty((N, N))
    bad = lib().crt_o_rmsd_cov_tm(aln, N, A, coords, offsets, r, c, t)
    return r, c, t, bad


def neighbor_joining(distance_matrix) -> Tuple[np.ndarray, np.ndarray]:
    """neighbor_joining.py:17-99: (tree uint64 [2N-3, 2], branch_lengths float64 [2N-3, 1])."""
    D = _c(distance_matrix)
    n = D.shape[0]
    if D.ndim != 2 or D.shape[1] != n:
        raise ValueError("square distance matrix expected")
    tree = np.zeros((max(2 * n - 3, 1), 2), np.uint64)
    bl = np.zeros(max(2 * n - 3, 1))
    k = int(lib().crt_o_neighbor_joining(D, n, tree, bl))
    if k < 0:
        raise IndexError("neighbor_joining needs at least 3 nodes (the reference indexes out of range)")
    return tree[:k], bl[:k].reshape(-1, 1)


def score_matrix(t1, c1, t2, c2, gamma_t=7.0, gamma_c=0.03, flexible=False) -> np.ndarray:
    """Protein.score_function, multiple_alignment.py:321-349: the full n x m matrix (flexible=True, :323-326: the tensor
    Gaussian alone)."""
    if flexible:
        return rbf_matrix(t1, t2, gamma_t)
    t1, c1, t2, c2 = _c(t1), _c(c1), _c(t2), _c(c2)
    S = np.empty((t1.shape[0], t2.shape[0]))
    lib().crt_o_score_matrix(t1, c1, t1.shape[0], t2, c2, t2.shape[0], t1.shape[1], gamma_t, gamma_c, S)
    return S


def mean_function(t1, c1, t2, c2, aln_1, aln_2, flexible=False):
    """Protein.mean_function, multiple_alignment.py:351-383 -> (tensors_mean, coordinates_mean); flexible=True (:359-360): the
    node has no coordinates (None)."""
    k = len(aln_1)
    tm = np.zeros((k, t1.shape[1]))
    for i, (x, y) in enumerate(zip(aln_1, aln_2)):
        tm[i] = t2[y] if x == -1 else (t1[x] if y == -1 else (t1[x] + t2[y]) / 2)
    if flexible:
        return tm, None
    p1, p2 = common_positions(aln_1, aln_2)
    if len(p1) <= 3:
        k1, k2 = np.array(c1), np.array(c2)
    else:
        k1, k2, _ = superpose_with_subset(c1, c2, c1[p1], c2[p2])
    cm = np.zeros((k, 3))
    for i, (x, y) in enumerate(zip(aln_1, aln_2)):
        cm[i] = k2[y] if x == -1 else (k1[x] if y == -1 else (k1[x] + k2[y]) / 2)
    return tm, cm


def mean_weights(w1, w2, aln_1, aln_2) -> np.ndarray:
    """get_mean_weights, multiple_alignment.py:73-82."""
    out = np.zeros((len(aln_1), 1))
    for i, (x, y) in enumerate(zip(aln_1, aln_2)):
        if x != -1:
            out[i] += w1[x]
        if y != -1:
            out[i] += w2[y]
    return out


def progressive_align(seqs, tree, gap_open=1.0, gap_extend=0.01, consensus_weight=1.0, gamma_weight=0.03,
                      gamma_t=7.0, gamma_c=0.03, want_final_alignments=False, flexible_score=False, flexible_mean=False):
    """MultipleAlignment.progressive_align, multiple_alignment.py:172-253, on [(name, tensors, coords)] (small cases:
    Python loops).  Returns (alignment {name: int64[A]}, final_sequences [(name, tensors, coords)], final_weights).
    flexible_score / flexible_mean: the flexible flags of score_function_params / mean_function_params."""
    fs = [(n, _c(t), None if c is None else _c(c)) for n, t, c in seqs]
    fa = {n: {n: np.arange(len(t))} for n, t, _ in fs}
    fw = [np.full((len(t), 1), consensus_weight, dtype=np.float64) for _, t, _ in fs]

    def node(n1, n2, n_int):
        (name_1, t1, c1), (name_2, t2, c2) = fs[n1], fs[n2]
        w1, w2 = fw[n1], fw[n2]
        l1, l2 = len(fa[name_1]), len(fa[name_2])
        mult1, mult2 = l2 / (2 * (l1 + l2)), l1 / (2 * (l1 + l2))
        name_int = f"int-{n_int}"
        S = score_matrix(t1, c1, t2, c2, gamma_t, gamma_c, flexible_score)
        S += rbf_matrix(w1 * mult1, w2 * mult2, gamma_weight)
        a1, a2, _ = dtw_align(S, gap_open, gap_extend)
        tmn, cmn = mean_function(t1, c1, t2, c2, a1, a2, flexible_mean)
        wmn = mean_weights(w1, w2, a1, a2)
        fa[name_1] = {k: np.array([v[i] if i != -1 else -1 for i in a1]) for k, v in fa[name_1].items()}
        fa[name_2] = {k: np.array([v[i] if i != -1 else -1 for i in a2]) for k, v in fa[name_2].items()}
        fa[name_int] = {**fa[name_1], **fa[name_2]}
        fs.append((name_int, tmn, cmn))
        fw.append(wmn)

    tree = np.asarray(tree)
    for x in range(0, tree.shape[0] - 1, 2):
        n1, n2, ni = int(tree[x, 0]), int(tree[x + 1, 0]), int(tree[x, 1])
        assert int(tree[x + 1, 1]) == ni
        node(n1, n2, ni)
    n1, n2 = int(tree[-1, 0]), int(tree[-1, 1])
    node(n1, n2, "final")
    alignment = {**fa[fs[n1][0]], **fa[fs[n2][0]]}
    if want_final_alignments:
        return alignment, fs, fw, fa
    return alignment, fs, fw


def progressive_node(t1, c1, w1, t2, c2, w2, mult1, mult2, gamma_t=7.0, gamma_c=0.03, gamma_weight=0.03, gap_open=1.0, gap_extend=0.01,
                     flexible_score=False, flexible_mean=False):
    """One make_intermediate_node (multiple_alignment.py:195-234) with the signature of Engine.progressive_node."""
    t1, c1, t2, c2 = _c(t1), _c(c1), _c(t2), _c(c2)
    w1, w2 = np.asarray(w1, dtype=np.float64).reshape(-1, 1), np.asarray(w2, dtype=np.float64).reshape(-1, 1)
    S = score_matrix(t1, c1, t2, c2, gamma_t, gamma_c, flexible_score)
    if not gamma_weight < 0:
        S = S + rbf_matrix(w1 * mult1, w2 * mult2, gamma_weight)
    a1, a2, sc = dtw_align(S, gap_open, gap_extend)
    tmn, cmn = mean_function(t1, c1, t2, c2, a1, a2, flexible_mean)
    return a1, a2, tmn, cmn, mean_weights(w1, w2, a1, a2), sc, 0


# --------------------------------------------------------------------------------------------------------------
# Consumers of the multiple alignment: control flow restated in Python over the C primitives.
# --------------------------------------------------------------------------------------------------------------
def coverage_gap_matrix(aln):
    """make_coverage_gap_distance_matrix, multiple_alignment.py:45-56."""
    aln = _c(aln, np.int64)
    n = aln.shape[0]
    dist, al = np.empty((n, n)), np.empty((n, n), np.int32)
    if lib().crt_o_coverage_gap_matrix(aln, n, aln.shape[1], dist, al) != 0:
        raise ZeroDivisionError("a protein has no residue in the alignment")
    return dist, al


def _reference_index(aln) -> int:
    """sorted(names, key=residues in the alignment, reverse=True)[0] -- the first protein with the maximum, :855."""
    return int(np.argmax((np.asarray(aln) != -1).sum(axis=1)))


def core_columns(aln) -> np.ndarray:
    """Columns where no protein has a gap, multiple_alignment.py:856-862."""
    return np.nonzero((np.asarray(aln) != -1).all(axis=0))[0]


def superpose_core(aln, coords_list, reference: int, core=None):
    """superpose_core, multiple_alignment.py:869-905.  Returns (new coordinate list, rotations, translations)."""
    aln = np.asarray(aln)
    core = core_columns(aln) if core is None else np.asarray(core)
    ref_core = np.ascontiguousarray(coords_list[reference][aln[reference][core]])
    cen = _mean_axis0(ref_core)
    ref_c = ref_core - cen
    out, rots, trans = [], [], []
    for i, c in enumerate(coords_list):
        if i == reference:
            out.append(c - cen); rots.append(np.eye(3)); trans.append(-cen)
            continue
        R, t = kabsch(ref_c, np.ascontiguousarray(c[aln[i][core]]))
        out.append(apply_rotran(c, R, t)); rots.append(R); trans.append(t)
    return out, np.array(rots), np.array(trans)


def _mean_axis0(x) -> np.ndarray:
    """helper.nb_mean_axis_0 (helper.py:45-53): np.mean of every column = sequential sum / n."""
    x = np.asarray(x, dtype=np.float64)
    out = np.zeros(x.shape[1])
    for a in range(x.shape[1]):
        acc = 0.0
        for v in x[:, a]:
            acc += float(v)
        out[a] = acc / x.shape[0]
    return out


def superpose_reference(aln, coords_list, reference: int):
    """superpose_reference, multiple_alignment.py:908-927 (the loop replaces the reference's own coordinates on its turn)."""
    aln = np.asarray(aln)
    cur = [np.array(c, dtype=np.float64) for c in coords_list]
    rots, trans, ncs = [], [], []
    for i in range(len(cur)):
        p1, p2 = common_positions(aln[reference], aln[i])
        ncs.append(len(p1))
        assert len(p1) > 3
        R, t = kabsch(np.ascontiguousarray(cur[reference][p1]), np.ascontiguousarray(cur[i][p2]))
        cur[i] = apply_rotran(cur[i], R, t)
        rots.append(R); trans.append(t)
    return cur, np.array(rots), np.array(trans), np.array(ncs)


def superpose(aln, coords_list):
    """superpose, multiple_alignment.py:854-867.  Returns (mode, reference, new coordinate list)."""
    aln = np.asarray(aln)
    ref = _reference_index(aln)
    core = core_columns(aln)
    if len(core) < aln.shape[1] // 2:
        return "reference", ref, superpose_reference(aln, coords_list, ref)[0]
    return "core", ref, superpose_core(aln, coords_list, ref, core)[0]


def get_reference_structures(aln, minimum_coverage=50):
    """get_reference_structures, multiple_alignment.py:740-784, on indices instead of names.
    Returns (first reference, {reference: [members]} in insertion order, [not aligning])."""
    aln = np.asarray(aln)
    n = aln.shape[0]
    dist, al = coverage_gap_matrix(aln)
    mincov = np.array([minimum_coverage * int((aln[i] != -1).sum()) / 100 for i in range(n)])
    refs = {}
    first = int(np.argmin(np.median(dist, axis=0)))
    not_cov = np.where(al[:, first] < mincov[:])[0]
    covered = list(np.where(al[:, first] >= mincov[:])[0])
    refs[first] = [int(c) for c in covered]
    problematic = []
    while len(not_cov) > 0:
        if len(not_cov) > 1:
            r = covered[int(np.argmin(np.median(dist[not_cov, :][:, covered], axis=0)))]
        else:
            r = covered[int(np.argmin(dist[not_cov, :][:, covered]))]
        cov_i = not_cov[np.where(al[not_cov, r] >= mincov[not_cov])[0]]
        if len(cov_i) == 0:
            problematic += list(not_cov)
            break
        not_cov = not_cov[np.where(al[not_cov, r] < mincov[not_cov])[0]]
        refs[int(r)] = [int(c) for c in cov_i]
        covered += list(cov_i)
    no_aligning = []
    for i in problematic:
        found = False
        for j in covered:
            if al[i, j] >= mincov[i]:
                refs[int(j)].append(int(i))
                found = True
                break
        if not found:
            no_aligning.append(int(i))
    return first, refs, no_aligning


def superpose_references(aln, coords_list, minimum_coverage=50):
    """superpose_references, multiple_alignment.py:930-950."""
    aln = np.asarray(aln)
    cur = [np.array(c, dtype=np.float64) for c in coords_list]
    first, refs, no_aligning = get_reference_structures(aln, minimum_coverage)
    for r, members in refs.items():
        for i in members:
            p1, p2 = common_positions(aln[r], aln[i])
            assert len(p1) > 3
            R, t = kabsch(np.ascontiguousarray(cur[r][p1]), np.ascontiguousarray(cur[i][p2]))
            cur[i] = apply_rotran(cur[i], R, t)
    return cur


def format_matrix(names, matrix) -> bytes:
    """The bytes helper.write_distance_matrix writes (helper.py:183-203)."""
    M = _c(np.asarray(matrix, dtype=np.float64)[:len(names)])
    enc = [str(x).encode("utf-8") for x in names]
    off = np.zeros(len(enc) + 1, np.int64)
    if enc:
        off[1:] = np.cumsum([len(e) for e in enc])
    blob = b"".join(enc)
    n = lib().crt_o_format_matrix(M, M.shape[0], M.shape[1], blob, off, None)
    buf = C.create_string_buffer(int(n) + 1)
    lib().crt_o_format_matrix(M, M.shape[0], M.shape[1], blob, off, C.cast(buf, C.c_void_p))
    return buf.raw[:n]


def format_fasta(names, sequences, aln) -> bytes:
    """MultipleAlignment.write_alignment, multiple_alignment.py:299-309."""
    out = []
    for name, seq, row in zip(names, sequences, np.asarray(aln)):
        out.append(">" + name + "\n" + "".join(seq[int(i)] if i != -1 else "-" for i in row) + "\n")
    return "".join(out).encode("utf-8")


def count_matrix(residues_list, alphabet_size: int) -> np.ndarray:
    """make_count_matrix, multiple_alignment.py:128-134."""
    arrs = [np.asarray(r, dtype=np.int64).reshape(-1) for r in residues_list]
    off = np.zeros(len(arrs) + 1, np.int64)
    off[1:] = np.cumsum([len(a) for a in arrs])
    idx = _c(np.concatenate(arrs), np.int64) if off[-1] else np.zeros(1, np.int64)
    out = np.empty((len(arrs), alphabet_size))
    if lib().crt_o_count_matrix(idx, off, len(arrs), alphabet_size, out) != 0:
        raise IndexError("shapemer index out of range")
    return out


def braycurtis(counts_1, counts_2) -> np.ndarray:
    """braycurtis, multiple_alignment.py:137-145."""
    a, b = _c(counts_1), _c(counts_2)
    out = np.empty((a.shape[0], b.shape[0]))
    lib().crt_o_braycurtis(a, a.shape[0], b, b.shape[0], a.shape[1], out)
    return out


def num_threads() -> int:
    return int(lib().crt_o_num_threads())
