"""ctypes binding of the C ABI in include/caretta_b200.h -- the only door between the Python host code and the
CUDA engine.  There is no CPU fallback: if the shared library is missing or no CUDA device is present, creating
an Engine raises."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CARETTA_B200_LIB") or os.path.join(_HERE, "libcaretta_b200.so")   # env override: A/B builds

FP64, FP32 = 0, 1
SUP_AUTO, SUP_CORE, SUP_REFERENCE = 0, 1, 2
ST_FEW_COMMON, ST_NO_POSITIVE, ST_NONFINITE, ST_TIE, ST_FP64 = 1, 2, 4, 8, 16
FLAG_FLEXIBLE = 1                                  # crt_params.flags: Protein.score_function(flexible=True)
# gamma_coords sentinels of progressive_node / progressive_level / msa_level (include/caretta_b200.h)
GC_FLEXIBLE, GC_FLEXIBLE_SCORE = -1.0, -2.0

EXPORTS = [
    "crt_last_error", "crt_version", "crt_create", "crt_destroy", "crt_device_info", "crt_set_chains", "crt_set_coords",
    "crt_pairwise_shard", "crt_shard_size", "crt_shard_pairs", "crt_plan_shard_size", "crt_plan_shard_pairs", "crt_fetch", "crt_fetch_device",
    "crt_last_elapsed_ms", "crt_last_phase_ms", "crt_last_launches", "crt_last_rerun", "crt_last_tc_pairs", "crt_last_cell_updates", "crt_last_traceback_bytes", "crt_pairwise_all", "crt_pairwise_list",
    "crt_sw_align_batch", "crt_dtw_align_batch", "crt_rmsd_cov_tm", "crt_rmsd_cov_tm_superposed", "crt_fp32_peak", "crt_host_alloc", "crt_host_free", "crt_neighbor_joining", "crt_progressive_node", "crt_progressive_level",
    "crt_score_matrix", "crt_mean_function", "crt_mean_weights",
    "crt_msa_begin", "crt_msa_level", "crt_msa_lengths", "crt_msa_fetch", "crt_msa_compose", "crt_msa_end",
    "crt_pack_results", "crt_scatter_gathered", "crt_multi_create", "crt_multi_destroy", "crt_multi_devices", "crt_multi_ctx",
    "crt_multi_set_chains", "crt_multi_pairwise_all", "crt_multi_last_timing",
    "crt_coverage_gap_matrix", "crt_superpose", "crt_superpose_pairs", "crt_format_matrix", "crt_format_fasta", "crt_text_fetch", "crt_count_matrix", "crt_braycurtis",
]


class CrtError(RuntimeError):
    pass


class Params(C.Structure):
    _fields_ = [("gamma_tensor", C.c_double), ("gamma_coords", C.c_double), ("sw_gap", C.c_double),
                ("precision", C.c_int32), ("flags", C.c_int32)]


_lib = None


def load_library():
    """Loads libcaretta_b200.so; raises if it has not been built (python -m caretta_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CrtError(f"{LIB_PATH} is missing: build it with `python -m caretta_b200.build` "
                       "(the engine has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    L.crt_last_error.restype = C.c_char_p
    L.crt_version.restype = C.c_int
    L.crt_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.crt_destroy.argtypes = [vp]
    L.crt_device_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i64)]
    L.crt_set_chains.argtypes = [vp, vp, vp, vp, i32, i32]
    L.crt_set_coords.argtypes = [vp, vp, vp, i32]
    L.crt_pairwise_shard.argtypes = [vp, C.POINTER(Params), i32, i32]
    L.crt_shard_size.argtypes = [vp, i32, i32]
    L.crt_shard_size.restype = i64
    L.crt_shard_pairs.argtypes = [vp, i32, i32, vp, vp]
    L.crt_plan_shard_size.argtypes = [vp, i32, i32, i32]
    L.crt_plan_shard_size.restype = i64
    L.crt_plan_shard_pairs.argtypes = [vp, i32, i32, i32, vp, vp]
    L.crt_fetch.argtypes = [vp, vp, vp, vp, vp, vp]
    L.crt_fetch_device.argtypes = [vp, vp, vp, vp, i64]
    L.crt_last_elapsed_ms.argtypes = [vp]
    L.crt_last_elapsed_ms.restype = dbl
    L.crt_last_phase_ms.argtypes = [vp, vp]
    L.crt_last_launches.argtypes = [vp]
    L.crt_last_launches.restype = i64
    L.crt_last_rerun.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_double)]
    L.crt_last_rerun.restype = C.c_int
    L.crt_last_tc_pairs.argtypes = [vp]
    L.crt_last_tc_pairs.restype = C.c_int64
    L.crt_last_cell_updates.argtypes = [vp]
    L.crt_last_cell_updates.restype = dbl
    L.crt_last_traceback_bytes.argtypes = [vp]
    L.crt_last_traceback_bytes.restype = dbl
    L.crt_pairwise_all.argtypes = [vp, C.POINTER(Params), vp, vp, vp]
    L.crt_pairwise_list.argtypes = [vp, C.POINTER(Params), vp, vp, i64, vp, vp, vp, vp, vp, vp, vp, vp, i64]
    L.crt_sw_align_batch.argtypes = [vp, vp, vp, vp, vp, i32, dbl, vp, vp, vp, i64, vp, vp]
    L.crt_dtw_align_batch.argtypes = [vp, vp, vp, vp, vp, i32, dbl, dbl, vp, vp, vp, i64, vp]
    L.crt_rmsd_cov_tm.argtypes = [vp, vp, i64, vp, vp, vp, C.POINTER(i32)]
    L.crt_rmsd_cov_tm_superposed.argtypes = [vp, vp, i64, vp, vp, vp, C.POINTER(i32)]
    L.crt_fp32_peak.argtypes = [vp, C.POINTER(dbl), C.POINTER(dbl)]
    L.crt_neighbor_joining.argtypes = [vp, vp, i32, vp, vp, C.POINTER(i64)]
    L.crt_progressive_node.argtypes = [vp, vp, vp, vp, i32, vp, vp, vp, i32, i32, dbl, dbl, dbl, dbl, dbl, dbl, dbl,
                                       vp, vp, C.POINTER(i32), vp, vp, vp, C.POINTER(dbl), C.POINTER(i32)]
    L.crt_score_matrix.argtypes = [vp, vp, vp, i32, vp, vp, i32, i32, dbl, dbl, i32, vp, C.POINTER(i32)]
    L.crt_mean_function.argtypes = [vp, vp, vp, i32, vp, vp, i32, i32, vp, vp, i64, i32, vp, vp, C.POINTER(i32)]
    L.crt_mean_weights.argtypes = [vp, vp, i32, vp, i32, vp, vp, i64, vp]
    L.crt_progressive_level.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp, dbl, dbl, dbl, dbl, dbl, vp, vp, vp, vp, vp, vp, vp, vp]
    L.crt_msa_begin.argtypes = [vp, dbl, C.POINTER(i32)]
    L.crt_msa_level.argtypes = [vp, i32, vp, vp, vp, dbl, dbl, dbl, dbl, dbl, vp, vp, i64, vp, vp, vp, vp, C.POINTER(i32)]
    L.crt_msa_lengths.argtypes = [vp, C.POINTER(i32), vp, i32]
    L.crt_msa_fetch.argtypes = [vp, vp, i32, vp, vp, vp]
    L.crt_msa_compose.argtypes = [vp, i32, vp, i32, C.POINTER(i32), vp, i64]
    L.crt_msa_end.argtypes = [vp]
    L.crt_coverage_gap_matrix.argtypes = [vp, vp, i32, i64, vp, vp]
    L.crt_superpose.argtypes = [vp, vp, i64, i32, i32, vp, i64, vp, vp, vp, vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i64)]
    L.crt_superpose_pairs.argtypes = [vp, vp, i64, vp, vp, i64, vp, i32, vp, vp, vp, vp]
    L.crt_format_matrix.argtypes = [vp, vp, i32, i32, vp, vp, C.POINTER(i64)]
    L.crt_format_fasta.argtypes = [vp, vp, i32, i64, vp, vp, vp, vp, C.POINTER(i64)]
    L.crt_text_fetch.argtypes = [vp, vp, i64]
    L.crt_count_matrix.argtypes = [vp, vp, vp, i32, i32, vp]
    L.crt_braycurtis.argtypes = [vp, vp, i32, vp, i32, i32, vp]
    L.crt_pack_results.argtypes = [vp, vp, i64, i32]
    L.crt_scatter_gathered.argtypes = [vp, vp, i32, i64, i32, vp, vp, vp]
    L.crt_multi_create.argtypes = [i32, vp, C.POINTER(vp)]
    L.crt_multi_destroy.argtypes = [vp]
    L.crt_multi_devices.argtypes = [vp]
    L.crt_multi_devices.restype = i32
    L.crt_multi_ctx.argtypes = [vp, i32]
    L.crt_multi_ctx.restype = vp
    L.crt_multi_set_chains.argtypes = [vp, vp, vp, vp, i32, i32]
    L.crt_multi_pairwise_all.argtypes = [vp, C.POINTER(Params), vp, vp, vp]
    L.crt_multi_last_timing.argtypes = [vp, vp, C.POINTER(i64)]
    L.crt_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    L.crt_host_free.argtypes = [vp]
    for name in EXPORTS:
        fn = getattr(L, name)
        if fn.restype is C.c_int and name not in ("crt_version",):
            fn.restype = C.c_int
    _lib = L
    return L


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class _PinnedBlock:
    """Owner of one crt_host_alloc block; freed when the last numpy view of it dies."""

    def __init__(self, nbytes: int):
        L = load_library()
        p = C.c_void_p()
        rc = L.crt_host_alloc(int(nbytes), C.byref(p))
        if rc != 0:
            raise CrtError(f"crt_host_alloc failed ({rc}): {L.crt_last_error().decode()}")
        self.ptr, self.nbytes, self._lib = p, int(nbytes), L

    def __del__(self):
        try:
            if self.ptr:
                self._lib.crt_host_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


def pinned_empty(shape, dtype=np.float64) -> np.ndarray:
    """numpy array in page-locked host memory (crt_host_alloc): host<->device copies of it run at full PCIe rate."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape))
    blk = _PinnedBlock(n * dtype.itemsize)
    buf = (C.c_char * max(blk.nbytes, 1)).from_address(blk.ptr.value)
    buf._crt_owner = blk            # array.base -> buf -> blk: the block lives as long as any view of the array
    return np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)


def pinned_like(a: np.ndarray) -> np.ndarray:
    out = pinned_empty(a.shape, a.dtype)
    out[...] = a
    return out


def plan_shard(offsets, rank: int, world: int):
    """Pairs (i, j) of one rank's shard, in result order, computed on the host (no device needed)."""
    L = load_library()
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    n = len(offsets) - 1
    np_ = int(L.crt_plan_shard_size(_p(offsets), n, rank, world))
    if np_ < 0:
        raise CrtError(f"crt_plan_shard_size failed: {L.crt_last_error().decode()}")
    pi = np.empty(max(np_, 1), np.int32)
    pj = np.empty(max(np_, 1), np.int32)
    rc = L.crt_plan_shard_pairs(_p(offsets), n, rank, world, _p(pi), _p(pj))
    if rc != 0:
        raise CrtError(f"crt_plan_shard_pairs failed: {L.crt_last_error().decode()}")
    return pi[:np_], pj[:np_]


class MultiEngine:
    """Every GPU of the box behind one call (crt_multi_*): one context per device, cost-sharded pairs, one NCCL all-gather inside
    the library, dense float64 matrices out.  devices: None = all visible devices."""

    def __init__(self, devices=None):
        self.lib = load_library()
        h = C.c_void_p()
        if devices is None:
            rc = self.lib.crt_multi_create(0, None, C.byref(h))
        else:
            ids = np.ascontiguousarray(list(devices), dtype=np.int32)
            rc = self.lib.crt_multi_create(len(ids), _p(ids), C.byref(h))
        if rc != 0:
            raise CrtError(f"crt_multi_create failed ({rc}): {self.lib.crt_last_error().decode()}")
        self.h = h
        self.n_devices = int(self.lib.crt_multi_devices(h))
        self.n_chains = 0
        self._offsets = None

    params = staticmethod(lambda *a, **k: Engine.params(*a, **k))

    def close(self):
        if getattr(self, "h", None):
            self.lib.crt_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise CrtError(f"{what} failed ({rc}): {self.lib.crt_last_error().decode()}")

    def set_chains(self, coords, tensors, offsets):
        coords = np.ascontiguousarray(coords, dtype=np.float64)
        tensors = np.ascontiguousarray(tensors, dtype=np.float64)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = len(offsets) - 1
        if coords.ndim != 2 or coords.shape[1] != 3 or tensors.ndim != 2 or coords.shape[0] != tensors.shape[0] \
                or offsets[-1] != coords.shape[0]:
            raise ValueError("coords [sumL,3], tensors [sumL,d], offsets [N+1] expected")
        self._check(self.lib.crt_multi_set_chains(self.h, _p(coords), _p(tensors), _p(offsets), n, tensors.shape[1]),
                    "crt_multi_set_chains")
        self.n_chains = n
        self._offsets = offsets.copy()

    def pairwise_all(self, prm: Params, want_rmsd_tm: bool = False, out=None):
        n = self.n_chains
        if out is not None:
            out = tuple(out) if isinstance(out, (tuple, list)) else (out,)
            for a in out:
                if a.dtype != np.float64 or a.shape != (n, n) or not a.flags.c_contiguous:
                    raise ValueError("out arrays must be C-contiguous float64 [N,N]")
            score = out[0]
            rm, tm = (out[1], out[2]) if want_rmsd_tm else (None, None)
        else:
            score = np.empty((n, n))
            rm = np.empty((n, n)) if want_rmsd_tm else None
            tm = np.empty((n, n)) if want_rmsd_tm else None
        self._check(self.lib.crt_multi_pairwise_all(self.h, C.byref(prm), _p(score), _p(rm), _p(tm)), "crt_multi_pairwise_all")
        return (score, rm, tm) if want_rmsd_tm else score

    def last_timing(self):
        out = np.zeros(3)
        n = C.c_int64(0)
        self._check(self.lib.crt_multi_last_timing(self.h, _p(out), C.byref(n)), "crt_multi_last_timing")
        return dict(shard_ms=float(out[0]), gather_ms=float(out[1]), wall_ms=float(out[2]), rerun_pairs=int(n.value))


class Engine:
    """One CUDA device.  Mirrors the life cycle the reference's driver has implicitly: hold the chains
    (MultipleAlignment.sequences), score all pairs (make_pairwise_matrix)."""

    def __init__(self, device: int = -1):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.crt_create(int(device), C.byref(h))
        if rc != 0:
            raise CrtError(f"crt_create failed ({rc}): {self.lib.crt_last_error().decode()}")
        self.h = h
        self.n_chains = 0
        self._offsets = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.crt_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise CrtError(f"{what} failed ({rc}): {self.lib.crt_last_error().decode()}")

    def device_info(self):
        sm, clk, mem = C.c_int32(), C.c_int32(), C.c_int64()
        self._check(self.lib.crt_device_info(self.h, C.byref(sm), C.byref(clk), C.byref(mem)), "crt_device_info")
        return dict(sm_count=sm.value, clock_khz=clk.value, mem_bytes=mem.value)

    @staticmethod
    def params(gamma_tensor=7.0, gamma_coords=0.03, precision=FP32, sw_gap=0.0, flexible=False) -> Params:
        return Params(float(gamma_tensor), float(gamma_coords), float(sw_gap), int(precision), FLAG_FLEXIBLE if flexible else 0)

    def set_chains(self, coords, tensors, offsets):
        coords = np.ascontiguousarray(coords, dtype=np.float64)
        tensors = np.ascontiguousarray(tensors, dtype=np.float64)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = len(offsets) - 1
        if coords.ndim != 2 or coords.shape[1] != 3 or tensors.ndim != 2 or coords.shape[0] != tensors.shape[0] \
                or offsets[-1] != coords.shape[0]:
            raise ValueError("coords [sumL,3], tensors [sumL,d], offsets [N+1] expected")
        self._check(self.lib.crt_set_chains(self.h, _p(coords), _p(tensors), _p(offsets), n, tensors.shape[1]),
                    "crt_set_chains")
        self.n_chains = n
        self._offsets = offsets.copy()
        self._tensor_width = int(tensors.shape[1])

    def set_coords(self, coords, offsets):
        """Chain set with coordinates only (enough for superpose* / rmsd_cov_tm; pair runs need set_chains)."""
        coords = np.ascontiguousarray(coords, dtype=np.float64)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        if coords.ndim != 2 or coords.shape[1] != 3 or offsets[-1] != coords.shape[0]:
            raise ValueError("coords [sumL,3] and offsets [N+1] expected")
        self._check(self.lib.crt_set_coords(self.h, _p(coords), _p(offsets), len(offsets) - 1), "crt_set_coords")
        self.n_chains = len(offsets) - 1
        self._offsets = offsets.copy()
        self._tensor_width = 1

    # ------------------------------------------------------------------------------------------------ all-vs-all
    def pairwise_all(self, prm: Params, want_rmsd_tm: bool = False, out=None):
        """Dense symmetric float64 [N,N] score matrix (and RMSD / TM by-products).  ``out``: optional preallocated
        C-contiguous float64 arrays (score,) or (score, rmsd, tm), e.g. from pinned_empty, filled in place."""
        n = self.n_chains
        if out is not None:
            out = tuple(out) if isinstance(out, (tuple, list)) else (out,)
            for a in out:
                if a.dtype != np.float64 or a.shape != (n, n) or not a.flags.c_contiguous:
                    raise ValueError("out arrays must be C-contiguous float64 [N,N]")
            score = out[0]
            rm, tm = (out[1], out[2]) if want_rmsd_tm else (None, None)
        else:
            score = np.empty((n, n))
            rm = np.empty((n, n)) if want_rmsd_tm else None
            tm = np.empty((n, n)) if want_rmsd_tm else None
        self._check(self.lib.crt_pairwise_all(self.h, C.byref(prm), _p(score), _p(rm), _p(tm)), "crt_pairwise_all")
        return (score, rm, tm) if want_rmsd_tm else score

    def pairwise_shard(self, prm: Params, rank: int = 0, world: int = 1):
        self._check(self.lib.crt_pairwise_shard(self.h, C.byref(prm), rank, world), "crt_pairwise_shard")

    def shard_size(self, rank: int, world: int) -> int:
        n = int(self.lib.crt_shard_size(self.h, rank, world))
        if n < 0:
            raise CrtError(f"crt_shard_size failed: {self.lib.crt_last_error().decode()}")
        return n

    def shard_pairs(self, rank: int, world: int):
        n = self.shard_size(rank, world)
        pi = np.empty(max(n, 1), np.int32)
        pj = np.empty(max(n, 1), np.int32)
        self._check(self.lib.crt_shard_pairs(self.h, rank, world, _p(pi), _p(pj)), "crt_shard_pairs")
        return pi[:n], pj[:n]

    def fetch(self, n: int, extras: bool = True):
        score = np.empty(max(n, 1))
        rm = np.empty(max(n, 1)) if extras else None
        tm = np.empty(max(n, 1)) if extras else None
        nc = np.empty(max(n, 1), np.int32) if extras else None
        st = np.empty(max(n, 1), np.int32) if extras else None
        self._check(self.lib.crt_fetch(self.h, _p(score), _p(rm), _p(tm), _p(nc), _p(st)), "crt_fetch")
        out = dict(score=score[:n])
        if extras:
            out.update(rmsd=rm[:n], tm=tm[:n], ncommon=nc[:n], status=st[:n])
        return out

    def pack_results(self, d_dst: int, pad: int, f64: bool = False):
        """score | rmsd | tm of the last shard packed into the DEVICE buffer d_dst [3 * pad] (float32, or float64 when f64)."""
        self._check(self.lib.crt_pack_results(self.h, C.c_void_p(d_dst), int(pad), 1 if f64 else 0), "crt_pack_results")

    def scatter_gathered(self, d_gathered: int, world: int, pad: int, f64: bool = False, want_rmsd_tm: bool = True, out=None):
        """[world][3][pad] all-gathered block (DEVICE address) -> dense symmetric float64 [N,N] score (, rmsd, tm) on the host.
        ``out``: optional preallocated C-contiguous float64 [N,N] arrays (score,) or (score, rmsd, tm)."""
        n = self.n_chains
        if out is not None:
            out = tuple(out) if isinstance(out, (tuple, list)) else (out,)
            for a in out:
                if a.dtype != np.float64 or a.shape != (n, n) or not a.flags.c_contiguous:
                    raise ValueError("out arrays must be C-contiguous float64 [N,N]")
            score = out[0]
            rm, tm = (out[1], out[2]) if want_rmsd_tm else (None, None)
        else:
            score = np.empty((n, n))
            rm = np.empty((n, n)) if want_rmsd_tm else None
            tm = np.empty((n, n)) if want_rmsd_tm else None
        self._check(self.lib.crt_scatter_gathered(self.h, C.c_void_p(d_gathered), int(world), int(pad), 1 if f64 else 0,
                                                  _p(score), _p(rm), _p(tm)), "crt_scatter_gathered")
        return (score, rm, tm) if want_rmsd_tm else score

    def fetch_device(self, d_score: int, d_rmsd: int, d_tm: int, n: int):
        self._check(self.lib.crt_fetch_device(self.h, C.c_void_p(d_score), C.c_void_p(d_rmsd), C.c_void_p(d_tm), n),
                    "crt_fetch_device")

    def last_elapsed_ms(self) -> float:
        return float(self.lib.crt_last_elapsed_ms(self.h))

    def last_phase_ms(self):
        out = np.zeros(4)
        self._check(self.lib.crt_last_phase_ms(self.h, _p(out)), "crt_last_phase_ms")
        return dict(fill1=out[0], trace=out[1], fill2=out[3])

    def last_traceback_bytes(self) -> float:
        return float(self.lib.crt_last_traceback_bytes(self.h))

    def last_launches(self) -> int:
        return int(self.lib.crt_last_launches(self.h))

    def last_rerun(self):
        """(pairs, device ms) the fp32 mode's tie detection sent through the float64 kernels in the last run."""
        n, ms = C.c_int64(0), C.c_double(0.0)
        self._check(self.lib.crt_last_rerun(self.h, C.byref(n), C.byref(ms)), "crt_last_rerun")
        return int(n.value), float(ms.value)

    def last_tc_pairs(self) -> int:
        """Pairs of the last run whose stage-1 fill ran on the tensor-core kernel (CARETTA_B200_TC=1)."""
        return int(self.lib.crt_last_tc_pairs(self.h))

    def last_cell_updates(self) -> float:
        return float(self.lib.crt_last_cell_updates(self.h))

    # ------------------------------------------------------------------------------------------------ pair list
    def pairwise_list(self, prm: Params, pi, pj, want_paths: bool = False):
        pi = np.ascontiguousarray(pi, dtype=np.int32)
        pj = np.ascontiguousarray(pj, dtype=np.int32)
        n = len(pi)
        score, rm, tm = np.empty(max(n, 1)), np.empty(max(n, 1)), np.empty(max(n, 1))
        nc, st = np.empty(max(n, 1), np.int32), np.empty(max(n, 1), np.int32)
        a1 = a2 = off = None
        cap = 0
        if want_paths:
            lens = np.diff(self._offsets)
            cap = int((lens[pi] + lens[pj]).sum()) + 1
            a1, a2 = np.empty(cap, np.int32), np.empty(cap, np.int32)
            off = np.zeros(n + 1, np.int64)
        self._check(self.lib.crt_pairwise_list(self.h, C.byref(prm), _p(pi), _p(pj), n, _p(score), _p(rm), _p(tm),
                                               _p(nc), _p(st), _p(a1), _p(a2), _p(off), cap), "crt_pairwise_list")
        out = dict(score=score[:n], rmsd=rm[:n], tm=tm[:n], ncommon=nc[:n], status=st[:n])
        if want_paths:
            out.update(aln1=a1, aln2=a2, aln_off=off)
        return out

    # ------------------------------------------------------------------------------------------------ DP in isolation
    @staticmethod
    def _pack_matrices(mats):
        mats = [np.ascontiguousarray(m, dtype=np.float64) for m in mats]
        for m in mats:
            if m.ndim != 2:
                raise ValueError("score matrices must be 2-D")
        n = np.array([m.shape[0] for m in mats], np.int32)
        mm = np.array([m.shape[1] for m in mats], np.int32)
        off = np.zeros(len(mats) + 1, np.int64)
        off[1:] = np.cumsum(n.astype(np.int64) * mm)
        flat = np.concatenate([m.ravel() for m in mats]) if mats else np.zeros(0)
        return flat, off, n, mm

    def _split(self, a1, a2, off, k):
        return [(a1[off[q]:off[q + 1]].astype(np.int64), a2[off[q]:off[q + 1]].astype(np.int64)) for q in range(k)]

    def dtw_align_batch(self, mats, gap_open: float, gap_extend: float):
        """[(aln1, aln2, score)] for each score matrix -- dtw.dtw_align (dynamic_time_warping.py:147-184)."""
        flat, off, n, m = self._pack_matrices(mats)
        k = len(mats)
        cap = int((n.astype(np.int64) + m + 1).sum()) + 1
        a1, a2 = np.empty(cap, np.int32), np.empty(cap, np.int32)
        aoff = np.zeros(k + 1, np.int64)
        score = np.empty(max(k, 1))
        self._check(self.lib.crt_dtw_align_batch(self.h, _p(flat), _p(off), _p(n), _p(m), k, float(gap_open),
                                                 float(gap_extend), _p(a1), _p(a2), _p(aoff), cap, _p(score)),
                    "crt_dtw_align_batch")
        return [(x, y, float(score[q])) for q, (x, y) in enumerate(self._split(a1, a2, aoff, k))]

    def sw_align_batch(self, mats, gap: float = 0.0, want_paths: bool = True):
        """[(aln1, aln2, score, status)] -- dtw.smith_waterman / smith_waterman_score (:204-278)."""
        flat, off, n, m = self._pack_matrices(mats)
        k = len(mats)
        cap = int((n.astype(np.int64) + m + 1).sum()) + 1
        a1 = np.empty(cap, np.int32) if want_paths else None
        a2 = np.empty(cap, np.int32) if want_paths else None
        aoff = np.zeros(k + 1, np.int64) if want_paths else None
        score = np.empty(max(k, 1))
        st = np.zeros(max(k, 1), np.int32)
        self._check(self.lib.crt_sw_align_batch(self.h, _p(flat), _p(off), _p(n), _p(m), k, float(gap), _p(a1), _p(a2),
                                                _p(aoff), cap, _p(score), _p(st)), "crt_sw_align_batch")
        if not want_paths:
            return [(None, None, float(score[q]), int(st[q])) for q in range(k)]
        return [(x, y, float(score[q]), int(st[q])) for q, (x, y) in enumerate(self._split(a1, a2, aoff, k))]

    def rmsd_cov_tm(self, aln, superpose: bool = True):
        """make_rmsd_coverage_tm_matrix on the chains of this engine; aln int64 [N, A].  superpose=True: Kabsch per pair
        (superpose_first=False, the reference's call site); False: the chains are already in one frame (after Engine.superpose)."""
        aln = np.ascontiguousarray(aln, dtype=np.int64)
        if aln.ndim != 2 or aln.shape[0] != self.n_chains:
            raise ValueError("aln must be [n_chains, A]")
        n = self.n_chains
        r, c, t = np.empty((n, n)), np.empty((n, n)), np.empty((n, n))
        bad = C.c_int32(0)
        fn = self.lib.crt_rmsd_cov_tm if superpose else self.lib.crt_rmsd_cov_tm_superposed
        self._check(fn(self.h, _p(aln), aln.shape[1], _p(r), _p(c), _p(t), C.byref(bad)), "crt_rmsd_cov_tm")
        return r, c, t, int(bad.value)

    def neighbor_joining(self, distance_matrix):
        """caretta/neighbor_joining.py:17-99 on the device: (tree uint64 [2N-3, 2], branch_lengths float64 [2N-3, 1])."""
        D = np.ascontiguousarray(distance_matrix, dtype=np.float64)
        if D.ndim != 2 or D.shape[0] != D.shape[1]:
            raise ValueError("square distance matrix expected")
        n = D.shape[0]
        if n < 3:
            raise IndexError("neighbor_joining needs at least 3 nodes (the reference indexes out of range below that)")
        tree = np.zeros((2 * n - 3, 2), np.uint64)
        bl = np.zeros((2 * n - 3, 1))
        k = C.c_int64()
        self._check(self.lib.crt_neighbor_joining(self.h, _p(D), n, _p(tree), _p(bl), C.byref(k)), "crt_neighbor_joining")
        return tree[:k.value], bl[:k.value]

    def progressive_node(self, t1, c1, w1, t2, c2, w2, mult1, mult2, gamma_tensor=7.0, gamma_coords=0.03, gamma_weight=0.03,
                         gap_open=1.0, gap_extend=0.01):
        """One node of progressive_align (multiple_alignment.py:195-234) on the device, float64.  Returns
        (aln_1 int64[k], aln_2 int64[k], tensors_mean [k,d], coordinates_mean [k,3], weights_mean [k,1], dtw_score, status)."""
        t1, c1, t2, c2 = (np.ascontiguousarray(a, dtype=np.float64) for a in (t1, c1, t2, c2))
        w1 = np.ascontiguousarray(np.asarray(w1, dtype=np.float64).reshape(-1))
        w2 = np.ascontiguousarray(np.asarray(w2, dtype=np.float64).reshape(-1))
        n, m, d = t1.shape[0], t2.shape[0], t1.shape[1]
        if c1.shape != (n, 3) or c2.shape != (m, 3) or t2.shape[1] != d or len(w1) != n or len(w2) != m:
            raise ValueError("tensors [L,d], coordinates [L,3], weights [L] expected for both sequences")
        cap = n + m + 1
        a1, a2 = np.empty(cap, np.int32), np.empty(cap, np.int32)
        tm, cm, wm = np.empty((cap, d)), np.empty((cap, 3)), np.empty(cap)
        k, sc, st = C.c_int32(), C.c_double(), C.c_int32()
        self._check(self.lib.crt_progressive_node(self.h, _p(t1), _p(c1), _p(w1), n, _p(t2), _p(c2), _p(w2), m, d,
                                                  float(mult1), float(mult2), float(gamma_tensor), float(gamma_coords),
                                                  float(gamma_weight), float(gap_open), float(gap_extend), _p(a1), _p(a2),
                                                  C.byref(k), _p(tm), _p(cm), _p(wm), C.byref(sc), C.byref(st)),
                    "crt_progressive_node")
        k = k.value
        return (a1[:k].astype(np.int64), a2[:k].astype(np.int64), tm[:k].copy(), cm[:k].copy(), wm[:k].reshape(-1, 1).copy(),
                sc.value, st.value)

    def score_matrix(self, t1, c1, t2, c2, gamma_tensor=0.03, gamma_coords=0.03, flexible=False):
        """Protein.score_function (multiple_alignment.py:321-349): (float64 [n,m] score matrix, status).  flexible=True: the tensor
        Gaussian alone, c1 / c2 may be None."""
        t1, t2 = (np.ascontiguousarray(a, dtype=np.float64) for a in (t1, t2))
        if t1.ndim != 2 or t2.ndim != 2 or t1.shape[1] != t2.shape[1]:
            raise ValueError("tensors [n,d] and [m,d] expected")
        n, m, d = t1.shape[0], t2.shape[0], t1.shape[1]
        c1, c2 = (None if a is None else np.ascontiguousarray(a, dtype=np.float64) for a in (c1, c2))
        if not flexible and (c1 is None or c2 is None or c1.shape != (n, 3) or c2.shape != (m, 3)):
            raise ValueError("coordinates [n,3] and [m,3] expected")
        if flexible and ((c1 is not None and c1.shape != (n, 3)) or (c2 is not None and c2.shape != (m, 3))):
            c1 = c2 = None
        S = np.empty((n, m))
        st = C.c_int32()
        self._check(self.lib.crt_score_matrix(self.h, _p(t1), _p(c1), n, _p(t2), _p(c2), m, d, float(gamma_tensor), float(gamma_coords),
                                              1 if flexible else 0, _p(S), C.byref(st)), "crt_score_matrix")
        return S, st.value

    def mean_function(self, t1, c1, t2, c2, aln_1, aln_2, flexible=False):
        """Protein.mean_function (multiple_alignment.py:351-383): (tensors_mean [k,d], coordinates_mean [k,3] or None, status)."""
        t1, t2 = (np.ascontiguousarray(a, dtype=np.float64) for a in (t1, t2))
        if t1.ndim != 2 or t2.ndim != 2 or t1.shape[1] != t2.shape[1]:
            raise ValueError("tensors [n,d] and [m,d] expected")
        n, m, d = t1.shape[0], t2.shape[0], t1.shape[1]
        a1, a2 = (np.ascontiguousarray(a, dtype=np.int64).reshape(-1) for a in (aln_1, aln_2))
        if len(a1) != len(a2):
            raise ValueError("aln_1 and aln_2 must have the same length")
        if not flexible:
            c1, c2 = (np.ascontiguousarray(a, dtype=np.float64) for a in (c1, c2))
            if c1.shape != (n, 3) or c2.shape != (m, 3):
                raise ValueError("coordinates [n,3] and [m,3] expected")
        else:
            c1 = c2 = None
        k = len(a1)
        tm = np.zeros((k, d))
        cm = None if flexible else np.zeros((k, 3))
        st = C.c_int32()
        self._check(self.lib.crt_mean_function(self.h, _p(t1), _p(c1), n, _p(t2), _p(c2), m, d, _p(a1), _p(a2), k, 1 if flexible else 0,
                                               _p(tm), _p(cm), C.byref(st)), "crt_mean_function")
        return tm, cm, st.value

    def mean_weights(self, w1, w2, aln_1, aln_2) -> np.ndarray:
        """get_mean_weights (multiple_alignment.py:73-82): float64 [k,1]."""
        w1 = np.ascontiguousarray(np.asarray(w1, dtype=np.float64).reshape(-1))
        w2 = np.ascontiguousarray(np.asarray(w2, dtype=np.float64).reshape(-1))
        a1, a2 = (np.ascontiguousarray(a, dtype=np.int64).reshape(-1) for a in (aln_1, aln_2))
        if len(a1) != len(a2):
            raise ValueError("aln_1 and aln_2 must have the same length")
        out = np.zeros((len(a1), 1))
        self._check(self.lib.crt_mean_weights(self.h, _p(w1), len(w1), _p(w2), len(w2), _p(a1), _p(a2), len(a1), _p(out)), "crt_mean_weights")
        return out

    def progressive_level(self, children, mults, gamma_tensor=7.0, gamma_coords=0.03, gamma_weight=0.03, gap_open=1.0, gap_extend=0.01):
        """All independent nodes of one guide-tree level in one device call (crt_progressive_level).  children: list of
        ((t1, c1, w1), (t2, c2, w2)) per node, mults: list of (multiplier_n1, multiplier_n2).  Returns one tuple per node, like
        progressive_node: (aln_1, aln_2, tensors_mean, coordinates_mean, weights_mean, dtw_score, status)."""
        k = len(children)
        if k == 0:
            return []
        seqs = [s for pair in children for s in pair]
        d = int(np.asarray(seqs[0][0]).shape[1])
        lens = np.array([np.asarray(s[0]).shape[0] for s in seqs], np.int64)
        off = np.zeros(2 * k + 1, np.int64)
        off[1:] = np.cumsum(lens)
        total = int(off[-1])
        T = np.concatenate([np.asarray(s[0], dtype=np.float64).reshape(-1, d) for s in seqs])
        X = np.concatenate([np.asarray(s[1], dtype=np.float64).reshape(-1, 3) for s in seqs])
        W = np.concatenate([np.asarray(s[2], dtype=np.float64).reshape(-1) for s in seqs])
        if T.shape != (total, d) or X.shape != (total, 3) or W.shape != (total,):
            raise ValueError("tensors [L,d], coordinates [L,3], weights [L] expected for every sequence")
        M = np.ascontiguousarray(np.asarray(mults, dtype=np.float64).reshape(k, 2))
        a1, a2 = np.empty(total, np.int32), np.empty(total, np.int32)
        ln = np.empty(k, np.int32)
        tm, cm, wm = np.empty((total, d)), np.empty((total, 3)), np.empty(total)
        sc, st = np.empty(k), np.empty(k, np.int32)
        self._check(self.lib.crt_progressive_level(self.h, k, d, _p(T), _p(X), _p(W), _p(off), _p(M), float(gamma_tensor), float(gamma_coords),
                                                   float(gamma_weight), float(gap_open), float(gap_extend), _p(a1), _p(a2), _p(ln), _p(tm),
                                                   _p(cm), _p(wm), _p(sc), _p(st)), "crt_progressive_level")
        out = []
        for q in range(k):
            lo, hi = int(off[2 * q]), int(off[2 * q]) + int(ln[q])
            # views of this call's freshly allocated arrays (nothing else refers to them)
            out.append((a1[lo:hi], a2[lo:hi], tm[lo:hi], cm[lo:hi], wm[lo:hi].reshape(-1, 1), float(sc[q]), int(st[q])))
        return out

    # ------------------------------------------------------------------------------------------------ device-resident MSA
    def msa_begin(self, consensus_weight: float) -> int:
        """Starts a progressive alignment on the chains of this engine: they become sequences 0..N-1 of the device pool."""
        # the nodes of the previous progressive alignment live in the pool this call replaces: whoever still holds a lazy view of
        # them (MultipleAlignment.final_sequences / final_consensus_weights) gets them fetched first, like the reference's lists
        prev = getattr(self, "_msa_outstanding", None)
        prev = prev() if prev is not None else None
        if prev is not None:
            prev.fetch()
        self._msa_outstanding = None
        n = C.c_int32()
        self._check(self.lib.crt_msa_begin(self.h, float(consensus_weight), C.byref(n)), "crt_msa_begin")
        self._msa_generation = getattr(self, "_msa_generation", 0) + 1
        self._msa_lengths = [int(x) for x in np.diff(self._offsets)]
        self._msa_d = self._tensor_width
        return n.value

    def msa_level(self, child1, child2, mults, gamma_tensor=7.0, gamma_coords=0.03, gamma_weight=0.03, gap_open=1.0, gap_extend=0.01):
        """All nodes (child1[k], child2[k]) of one tree level on the pool (crt_msa_level).  Returns (first_new_id,
        [(aln_1 int32 view, aln_2 int32 view, dtw_score, status)])."""
        c1 = np.ascontiguousarray(child1, dtype=np.int32)
        c2 = np.ascontiguousarray(child2, dtype=np.int32)
        k = len(c1)
        M = np.ascontiguousarray(np.asarray(mults, dtype=np.float64).reshape(k, 2))
        L = self._msa_lengths
        cap = int(sum(L[a] + L[b] for a, b in zip(c1.tolist(), c2.tolist())))
        a1, a2 = np.empty(max(cap, 1), np.int32), np.empty(max(cap, 1), np.int32)
        off, ln = np.zeros(k + 1, np.int64), np.empty(k, np.int32)
        sc, st = np.empty(k), np.empty(k, np.int32)
        first = C.c_int32()
        self._check(self.lib.crt_msa_level(self.h, k, _p(c1), _p(c2), _p(M), float(gamma_tensor), float(gamma_coords), float(gamma_weight),
                                           float(gap_open), float(gap_extend), _p(a1), _p(a2), cap, _p(off), _p(ln), _p(sc), _p(st),
                                           C.byref(first)), "crt_msa_level")
        self._msa_lengths.extend(int(x) for x in ln)
        out = [(a1[int(off[q]):int(off[q]) + int(ln[q])], a2[int(off[q]):int(off[q]) + int(ln[q])], float(sc[q]), int(st[q])) for q in range(k)]
        return first.value, out

    def msa_level_ext(self, child1, child2, mults, gamma_tensor=7.0, gamma_coords=0.03, gamma_weight=0.03, gap_open=1.0, gap_extend=0.01):
        """msa_level whose alignments are views that end in a -1 sentinel (entry [len] = -1: a gap index -1 then gathers -1):
        [(aln_1 int32 [len + 1], aln_2 int32 [len + 1], dtw_score, status)], no per-node copies."""
        c1 = np.ascontiguousarray(child1, dtype=np.int32)
        c2 = np.ascontiguousarray(child2, dtype=np.int32)
        k = len(c1)
        M = np.ascontiguousarray(np.asarray(mults, dtype=np.float64).reshape(k, 2))
        L = np.asarray(self._msa_lengths, dtype=np.int64)
        caps = L[c1] + L[c2]
        cap = int(caps.sum())
        # one spare slot at the end: the sentinel of a last node that fills its whole capacity (all-gap alignments only)
        a1, a2 = np.empty(cap + 1, np.int32), np.empty(cap + 1, np.int32)
        off, ln = np.zeros(k + 1, np.int64), np.empty(k, np.int32)
        sc, st = np.empty(k), np.empty(k, np.int32)
        first = C.c_int32()
        self._check(self.lib.crt_msa_level(self.h, k, _p(c1), _p(c2), _p(M), float(gamma_tensor), float(gamma_coords), float(gamma_weight),
                                           float(gap_open), float(gap_extend), _p(a1), _p(a2), cap, _p(off), _p(ln), _p(sc), _p(st),
                                           C.byref(first)), "crt_msa_level")
        self._msa_lengths.extend(ln.tolist())
        ends = off[:-1] + ln
        room = ln < caps                         # the sentinel fits into the node's own slots ...
        room[k - 1] = True                       # ... or into the spare slot behind the last node
        a1[ends[room]] = -1
        a2[ends[room]] = -1
        lo, hi, scl, stl = off[:-1].tolist(), (ends + 1).tolist(), sc.tolist(), st.tolist()
        out = [(a1[lo[q]:hi[q]], a2[lo[q]:hi[q]], scl[q], stl[q]) for q in range(k)]
        for q in np.nonzero(~room)[0].tolist():  # a node that fills its whole capacity (no column pairs two residues): copies
            out[q] = (np.append(a1[lo[q]:hi[q] - 1], np.int32(-1)), np.append(a2[lo[q]:hi[q] - 1], np.int32(-1)), scl[q], stl[q])
        return first.value, out

    def msa_fetch(self, ids):
        """[(tensors [L,d], coordinates [L,3], weights [L,1])] of the listed pool sequences (views of three packed arrays)."""
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        if len(ids) == 0:
            return []
        L = self._msa_lengths
        lens = np.array([L[i] for i in ids.tolist()], np.int64)
        off = np.concatenate([[0], np.cumsum(lens)])
        rows = int(off[-1])
        dd = self._msa_d
        T, X, W = np.empty((rows, dd)), np.empty((rows, 3)), np.empty(rows)
        self._check(self.lib.crt_msa_fetch(self.h, _p(ids), len(ids), _p(T), _p(X), _p(W)), "crt_msa_fetch")
        return [(T[off[q]:off[q + 1]], X[off[q]:off[q + 1]], W[off[q]:off[q + 1]].reshape(-1, 1)) for q in range(len(ids))]

    def msa_compose(self, root: int):
        """(pool ids of the sequences under `root` in dictionary order, int64 [n_under, A] index arrays in root's frame): crt_msa_compose."""
        A = int(self._msa_lengths[root])
        cap = len(self._offsets) - 1
        ids = np.empty(cap, np.int32)
        out = np.empty((cap, A), np.int64)
        n = C.c_int32()
        self._check(self.lib.crt_msa_compose(self.h, int(root), _p(ids), cap, C.byref(n), _p(out), out.size), "crt_msa_compose")
        return ids[:n.value], out[:n.value]

    def msa_track(self, nodes):
        """Registers the lazy view of the pool's nodes (weakly): it is fetched before the pool is replaced or released."""
        import weakref
        self._msa_outstanding = weakref.ref(nodes)

    def msa_end(self):
        self._check(self.lib.crt_msa_end(self.h), "crt_msa_end")
        self._msa_outstanding = None

    # ------------------------------------------------------------------------------------------------ alignment consumers
    def _aln(self, aln, need_chains=True):
        aln = np.ascontiguousarray(aln, dtype=np.int64)
        if aln.ndim != 2 or (need_chains and aln.shape[0] != self.n_chains):
            raise ValueError("aln must be int64 [n_chains, A]")
        return aln

    def coverage_gap_matrix(self, aln):
        """make_coverage_gap_distance_matrix (multiple_alignment.py:45-56): (distance float64 [N,N], aligning int32 [N,N])."""
        aln = self._aln(aln, need_chains=False)
        n = aln.shape[0]
        dist, al = np.empty((n, n)), np.empty((n, n), np.int32)
        self._check(self.lib.crt_coverage_gap_matrix(self.h, _p(aln), n, aln.shape[1], _p(dist), _p(al)), "crt_coverage_gap_matrix")
        return dist, al

    def superpose(self, aln, mode: int = SUP_AUTO, reference: int = -1, core_columns=None):
        """superpose / superpose_core / superpose_reference (multiple_alignment.py:854-927) on the chains of this engine.
        Returns dict(coords [sumL,3], rot [N,3,3], tran [N,3], ncommon [N], mode, reference, n_core)."""
        aln = self._aln(aln)
        n = self.n_chains
        coords = np.empty((int(self._offsets[-1]), 3))
        rot, tran, nc = np.empty((n, 3, 3)), np.empty((n, 3)), np.empty(n, np.int32)
        m, r, k = C.c_int32(), C.c_int32(), C.c_int64()
        cc = None if core_columns is None else np.ascontiguousarray(core_columns, dtype=np.int64)
        if cc is not None and len(cc) == 0:
            raise IndexError("empty core_indices")
        self._check(self.lib.crt_superpose(self.h, _p(aln), aln.shape[1], int(mode), int(reference), _p(cc), 0 if cc is None else len(cc),
                                           _p(coords), _p(rot), _p(tran), _p(nc), C.byref(m), C.byref(r), C.byref(k)), "crt_superpose")
        return dict(coords=coords, rot=rot, tran=tran, ncommon=nc, mode=m.value, reference=r.value, n_core=k.value)

    def superpose_pairs(self, aln, ref, mem, batch_off):
        """Sequential batches of (reference, member) superpositions over common alignment columns (superpose_references,
        multiple_alignment.py:930-950).  Returns dict(coords, rot [P,3,3], tran [P,3], ncommon [P])."""
        aln = self._aln(aln)
        ref = np.ascontiguousarray(ref, dtype=np.int32)
        mem = np.ascontiguousarray(mem, dtype=np.int32)
        batch_off = np.ascontiguousarray(batch_off, dtype=np.int64)
        k = len(ref)
        if len(mem) != k or len(batch_off) < 1:
            raise ValueError("ref / mem of equal length and batch_off [n_batches + 1] expected")
        coords = np.empty((int(self._offsets[-1]), 3))
        rot, tran, nc = np.empty((max(k, 1), 3, 3)), np.empty((max(k, 1), 3)), np.empty(max(k, 1), np.int32)
        self._check(self.lib.crt_superpose_pairs(self.h, _p(aln), aln.shape[1], _p(ref), _p(mem), k, _p(batch_off), len(batch_off) - 1,
                                                 _p(coords), _p(rot), _p(tran), _p(nc)), "crt_superpose_pairs")
        return dict(coords=coords, rot=rot[:k], tran=tran[:k], ncommon=nc[:k])

    @staticmethod
    def _pack_text(items):
        enc = [x if isinstance(x, bytes) else str(x).encode("utf-8") for x in items]
        off = np.zeros(len(enc) + 1, np.int64)
        if enc:
            off[1:] = np.cumsum([len(e) for e in enc])
        blob = np.frombuffer(b"".join(enc), dtype=np.uint8).copy() if off[-1] else np.zeros(1, np.uint8)
        return blob, off

    def _fetch_text_view(self, n: int) -> memoryview:
        """The text of the last crt_format_* call in a page-locked staging buffer the engine keeps (grow-only): valid until the
        next call.  Large texts (a 5000 x 5000 matrix is 198 MB) go from here straight to the file, without another copy."""
        buf = getattr(self, "_text_staging", None)
        if buf is None or buf.size < n:
            self._text_staging = buf = pinned_empty(max(n + n // 8, 1 << 16), np.uint8)
        self._check(self.lib.crt_text_fetch(self.h, _p(buf), n), "crt_text_fetch")
        return memoryview(buf)[:n]

    def _fetch_text(self, n: int) -> bytes:
        return bytes(self._fetch_text_view(n))

    def format_matrix(self, names, matrix, view: bool = False):
        """The bytes helper.write_distance_matrix (helper.py:183-203) writes: header, then 'name v v ...' rows with %.4f values.
        view=True: a memoryview of the engine's page-locked staging buffer (valid until the next format call) instead of bytes."""
        M = np.ascontiguousarray(matrix, dtype=np.float64)
        if M.ndim != 2 or M.shape[0] < len(names):
            raise IndexError("distance_matrix needs one row per name")
        M = np.ascontiguousarray(M[:len(names)])
        blob, off = self._pack_text(names)
        n = C.c_int64()
        self._check(self.lib.crt_format_matrix(self.h, _p(M), M.shape[0], M.shape[1], _p(blob), _p(off), C.byref(n)), "crt_format_matrix")
        return self._fetch_text_view(n.value) if view else self._fetch_text(n.value)

    def format_fasta(self, names, sequences, aln, view: bool = False):
        """The bytes MultipleAlignment.write_alignment (multiple_alignment.py:299-309) writes for aln int64 [N, A]."""
        aln = self._aln(aln, need_chains=False)
        if len(names) != aln.shape[0] or len(sequences) != aln.shape[0]:
            raise ValueError("one name and one sequence per alignment row expected")
        nb, noff = self._pack_text(names)
        sb, soff = self._pack_text(sequences)
        n = C.c_int64()
        self._check(self.lib.crt_format_fasta(self.h, _p(aln), aln.shape[0], aln.shape[1], _p(sb), _p(soff), _p(nb), _p(noff), C.byref(n)),
                    "crt_format_fasta")
        return self._fetch_text_view(n.value) if view else self._fetch_text(n.value)

    def count_matrix(self, residues_list, alphabet_size: int) -> np.ndarray:
        """make_count_matrix (multiple_alignment.py:128-134): float64 [N, alphabet_size] shapemer counts."""
        arrs = [np.asarray(r, dtype=np.int64).reshape(-1) for r in residues_list]
        off = np.zeros(len(arrs) + 1, np.int64)
        if arrs:
            off[1:] = np.cumsum([len(a) for a in arrs])
        idx = np.ascontiguousarray(np.concatenate(arrs)) if off[-1] else np.zeros(1, np.int64)
        out = np.empty((len(arrs), int(alphabet_size)))
        self._check(self.lib.crt_count_matrix(self.h, _p(idx), _p(off), len(arrs), int(alphabet_size), _p(out)), "crt_count_matrix")
        return out

    def braycurtis(self, counts_1, counts_2) -> np.ndarray:
        """braycurtis (multiple_alignment.py:137-145): float64 [n1, n2]."""
        a = np.ascontiguousarray(counts_1, dtype=np.float64)
        b = a if counts_2 is counts_1 else np.ascontiguousarray(counts_2, dtype=np.float64)
        if a.ndim != 2 or b.ndim != 2 or a.shape[1] != b.shape[1]:
            raise ValueError("count matrices [n1, K] and [n2, K] expected")
        out = np.empty((a.shape[0], b.shape[0]))
        self._check(self.lib.crt_braycurtis(self.h, _p(a), a.shape[0], _p(b), b.shape[0], a.shape[1], _p(out)), "crt_braycurtis")
        return out

    def fp32_peak(self):
        v, ms = C.c_double(), C.c_double()
        self._check(self.lib.crt_fp32_peak(self.h, C.byref(v), C.byref(ms)), "crt_fp32_peak")
        return v.value, ms.value
