"""Host-side mirror of the reference's interface for the all-vs-all pair path (names, argument meaning and error
behaviour follow caretta/multiple_alignment.py of TurtleTools/caretta 0.2.0), backed by the CUDA engine through
the C ABI (caretta_b200.engine).  Nothing here computes the path on the CPU.

    reference                                                        here
    ---------------------------------------------------------------  --------------------------------------------
    Protein(name, tensors, coordinates, sequence)        :312-319    Protein (same fields, same __len__/__str__)
    MultipleAlignment(sequences)                         :148-156    MultipleAlignment
      .make_pairwise_matrix(score_function_params)       :158-170    same signature, float64 [N,N], symmetric, diag 0
    make_rmsd_coverage_tm_matrix(alignment, proteins,    :1000-1055  same signature
                                 superpose_first)
    make_coverage_gap_distance_matrix(alignment_array)   :45-56      same signature
    get_reference_structures(alignment, min_coverage)    :740-784    same signature
    superpose / superpose_core / superpose_reference /   :854-950    same signatures (coordinates of the given proteins are
      superpose_references                                           replaced, the list is returned)
    MultipleAlignment.to_sequence_alignment /            :287-309    same signatures, same bytes
      write_alignment
    helper.write_distance_matrix(names, matrix, file)    helper.py   write_distance_matrix, same bytes
                                                         :183-203
    make_count_matrix / braycurtis (fast-mode guide)     :128-145    same signatures, bit-identical
    dtw.dtw_align / smith_waterman / smith_waterman_score            dtw_align / smith_waterman / smith_waterman_score
      (dynamic_time_warping.py:147-278)                              (batched: *_batch)
    StructureMultiple (name used by the CLI help and the legacy API) StructureMultiple facade

``install(reference_module)`` swaps the reference's own methods for these, so that an unmodified
``align_from_structure_files`` / ``caretta-cli`` run uses the GPU for the pair loop.
"""
from __future__ import annotations

import abc
import collections.abc
import os
import typing
from dataclasses import dataclass, field

import numpy as np

from . import engine as _engine

DEFAULT_SCORE_PARAMS = dict(flexible=False, gamma_tensor=7.0, gamma_coords=0.03)     # multiple_alignment.py:490-492


def _precision_from_env() -> int:
    v = os.environ.get("CARETTA_B200_PRECISION", "fp32").lower()
    if v in ("fp64", "f64", "double", "parity"):
        return _engine.FP64
    if v in ("fp32", "f32", "float"):
        return _engine.FP32
    raise ValueError(f"CARETTA_B200_PRECISION={v!r}: expected fp32 or fp64")


class SequenceBase(abc.ABC):
    """The type the reference's driver is generic over (multiple_alignment.py:109-127): anything with a score_function, a
    mean_function, a length and a string form.  Sequences that carry .tensors (and .coordinates) take the fused device path;
    any other SequenceBase goes through its own score_function / mean_function with the DPs on the device."""
    name: str

    @abc.abstractmethod
    def score_function(self, other: "SequenceBase", **kwargs) -> np.ndarray:
        pass

    def mean_function(self, other: "SequenceBase", aln_1: np.ndarray, aln_2: np.ndarray, name_int: str, **kwargs) -> "SequenceBase":
        pass

    @abc.abstractmethod
    def __len__(self) -> int:
        pass

    @abc.abstractmethod
    def __str__(self) -> str:
        pass


def _echo_few_positions(name_1, name_2) -> None:
    print(f"Too few aligning positions for {name_1} and {name_2}, continuing without superposition")


@dataclass
class Protein(SequenceBase):
    """Same fields and methods as the reference's Protein (multiple_alignment.py:312-389); the methods run on the device."""
    name: str
    tensors: np.ndarray
    coordinates: np.ndarray = None
    sequence: str = ""

    def score_function(self, other: "Protein", flexible=False, gamma_tensor=0.03, gamma_coords=0.03, verbose=True) -> np.ndarray:
        """multiple_alignment.py:321-349: the float64 [len(self), len(other)] score matrix (crt_score_matrix)."""
        S, status = get_engine().score_matrix(self.tensors, None if flexible else self.coordinates, other.tensors,
                                              None if flexible else other.coordinates, gamma_tensor, gamma_coords, flexible)
        if verbose and status & _engine.ST_FEW_COMMON:
            _echo_few_positions(self.name, other.name)
        return S

    def mean_function(self, other: "Protein", aln_1: np.ndarray, aln_2: np.ndarray, name_int: str, flexible=False,
                      verbose=True) -> "Protein":
        """multiple_alignment.py:351-383: the intermediate node of an alignment of self and other (crt_mean_function)."""
        tm, cm, status = get_engine().mean_function(self.tensors, None if flexible else self.coordinates, other.tensors,
                                                    None if flexible else other.coordinates, aln_1, aln_2, flexible)
        if verbose and status & _engine.ST_FEW_COMMON:
            _echo_few_positions(self.name, other.name)
        return Protein(name_int, tm) if flexible else Protein(name_int, tm, cm)

    def __len__(self) -> int:
        return self.tensors.shape[0]

    def __str__(self) -> str:
        return self.sequence


def get_mean_weights(weights_1: np.ndarray, weights_2: np.ndarray, aln_1: np.ndarray, aln_2: np.ndarray) -> np.ndarray:
    """multiple_alignment.py:73-82: float64 [len(aln_1), 1] (crt_mean_weights)."""
    return get_engine().mean_weights(weights_1, weights_2, aln_1, aln_2)


def _on_fused_path(sequences) -> bool:
    """True when every sequence carries shape tensors: the whole pair recipe then runs inside the device kernels."""
    return all(getattr(s, "tensors", None) is not None for s in sequences)


def pack_sequences(sequences, need_coordinates: bool = True, staging=None) -> typing.Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """List of Protein-like objects (anything with .tensors [L,d] and .coordinates [L,3]) -> packed chain set.
    need_coordinates=False (flexible=True runs, where the reference never reads them, multiple_alignment.py:323-326): a sequence
    whose coordinates are None gets zeros."""
    if len(sequences) == 0:
        raise ValueError("no sequences")
    asarray = np.asarray
    ts = [asarray(s.tensors) for s in sequences]
    d = int(ts[0].shape[1]) if ts[0].ndim == 2 else -1
    cs = []
    for s, t in zip(sequences, ts):
        if t.ndim != 2 or t.shape[1] != d:
            raise ValueError(f"{getattr(s, 'name', '?')}: tensors must be [L,{d}]")
        c = getattr(s, "coordinates", None)
        if c is None and not need_coordinates:
            c = np.zeros((t.shape[0], 3))
        else:
            c = asarray(c)
        if c.ndim != 2 or c.shape[1] != 3 or c.shape[0] != t.shape[0]:
            raise ValueError(f"{getattr(s, 'name', '?')}: coordinates must be [L,3] with the same L as tensors")
        cs.append(c)
    offsets = np.zeros(len(sequences) + 1, np.int64)
    np.cumsum([t.shape[0] for t in ts], out=offsets[1:])
    if staging is not None:
        coords, tensors = staging.views(int(offsets[-1]), d)
        np.concatenate(cs, out=coords)
        np.concatenate(ts, out=tensors)
        return coords, tensors, offsets
    coords = np.concatenate(cs, dtype=np.float64)
    tensors = np.concatenate(ts, dtype=np.float64)
    return coords, tensors, offsets


class _PinnedStaging:
    """Page-locked staging buffers for the packed chain set of the mirror's own engine calls: set_chains copies from them at the
    full PCIe rate instead of through the driver's bounce buffers (a third of progressive_align's host time at N = 1000), and
    returns only when the copy is done, so the next call may overwrite them.  They grow and are never handed to the caller."""

    def __init__(self):
        self.coords = self.tensors = None

    def views(self, rows: int, d: int):
        if self.coords is None or self.coords.size < rows * 3:
            self.coords = _engine.pinned_empty(max(rows * 3 + rows * 3 // 4, 1))
        if self.tensors is None or self.tensors.size < rows * d:
            self.tensors = _engine.pinned_empty(max(rows * d + rows * d // 4, 1))
        return self.coords[:rows * 3].reshape(rows, 3), self.tensors[:rows * d].reshape(rows, d)


_staging: typing.Optional[_PinnedStaging] = None


def _pack_staged(sequences, need_coordinates: bool = True):
    """pack_sequences into the process's pinned staging buffers (for an immediate set_chains); plain arrays if pinning fails or
    CARETTA_B200_PINNED_STAGING=0."""
    global _staging
    if os.environ.get("CARETTA_B200_PINNED_STAGING", "1") == "0":
        return pack_sequences(sequences, need_coordinates)
    if _staging is None:
        _staging = _PinnedStaging()
    try:
        return pack_sequences(sequences, need_coordinates, staging=_staging)
    except _engine.CrtError:
        return pack_sequences(sequences, need_coordinates)


def pack_coordinates(sequences) -> typing.Tuple[np.ndarray, np.ndarray]:
    """Coordinates and offsets only (what the consumers of an alignment need)."""
    if len(sequences) == 0:
        raise ValueError("no sequences")
    lens = []
    for s in sequences:
        c = np.asarray(s.coordinates)
        if c.ndim != 2 or c.shape[1] != 3 or c.shape[0] == 0:
            raise ValueError(f"{getattr(s, 'name', '?')}: coordinates must be [L,3]")
        lens.append(c.shape[0])
    offsets = np.zeros(len(sequences) + 1, np.int64)
    offsets[1:] = np.cumsum(lens)
    return np.concatenate([np.asarray(s.coordinates, dtype=np.float64) for s in sequences]), offsets


_shared_engine: typing.Optional[_engine.Engine] = None


def get_engine() -> _engine.Engine:
    """One engine (context) per process, on the current CUDA device (LOCAL_RANK when launched by torchrun)."""
    global _shared_engine
    if _shared_engine is None:
        dev = int(os.environ.get("CARETTA_B200_DEVICE", os.environ.get("LOCAL_RANK", "-1")))
        _shared_engine = _engine.Engine(dev)
    return _shared_engine


_shared_multi: typing.Optional[_engine.MultiEngine] = None
# below this many residue pairs a second GPU costs more (upload + all-gather) than it saves
MULTI_GPU_MIN_CELLS = 2.0e9


_multi_probe: typing.Optional[str] = None


def get_multi_engine() -> typing.Optional[_engine.MultiEngine]:
    """Every visible GPU behind the one call the reference makes (multiple_alignment.py:498-500): crt_multi_* (one context per
    device, NCCL all-gather inside the library).  CARETTA_B200_DEVICES = "all" (default), "1" / "0,2,3" (a list of device ids);
    a single id, a torchrun launch (LOCAL_RANK set: one process per GPU, caretta_b200.distributed) or a one-GPU box -> None."""
    global _shared_multi, _multi_probe
    if _shared_multi is not None:
        return _shared_multi
    spec = os.environ.get("CARETTA_B200_DEVICES", "all").strip().lower()
    if "LOCAL_RANK" in os.environ or "CARETTA_B200_DEVICE" in os.environ:
        return None
    if _multi_probe == spec:                   # this device list was tried before and gave fewer than two devices:
        return None                            # creating and destroying a context set costs 10 ms and more per call
    devices = None if spec in ("all", "") else [int(x) for x in spec.split(",") if x.strip() != ""]
    if devices is not None and len(devices) < 2:
        return None
    try:
        m = _engine.MultiEngine(devices)
    except _engine.CrtError:
        _multi_probe = spec
        return None
    if m.n_devices < 2:
        m.close()
        _multi_probe = spec
        return None
    _shared_multi = m
    return m


class _NodeAlignments(collections.abc.Mapping):
    """MultipleAlignment.final_alignments of the reference (multiple_alignment.py:181-183, :219-232): node name -> {member name ->
    int64 index array}, same keys and order.  A read-only mapping; a node's dictionary is composed from the stored pairwise
    alignments when it is first read (at N = 5000 the eager dictionaries are 140 000 arrays nobody may ever read)."""

    def __init__(self, names, build):
        self._index = {n: i for i, n in enumerate(names)}
        self._names, self._build = list(names), build
        self._cache = {}

    def __getitem__(self, name):
        i = self._index[name]
        d = self._cache.get(i)
        if d is None:
            d = self._cache[i] = self._build(i)
        return d

    def __iter__(self):
        return iter(self._names)

    def __len__(self):
        return len(self._names)

    def __reduce__(self):
        # pickled (the reference's --write-class dumps the whole MultipleAlignment, :557-559) and deep-copied as the plain
        # dictionary of dictionaries the reference holds
        return dict, ({name: dict(self[name]) for name in self._names},)


class _PoolNodes:
    """The intermediate nodes of a progressive alignment that ran on the device pool (crt_msa_*): fetched from the device in one
    call the first time any of them is read -- or, when nobody has read them yet, right before the next progressive alignment on
    the same engine replaces the pool (Engine.msa_begin fetches the outstanding view), so they stay readable for as long as the
    MultipleAlignment lives, like the reference's lists (multiple_alignment.py:251-252).  The pool's device buffers are kept for
    the next alignment (cudaMalloc / cudaFree of them cost more than an alignment); get_engine().msa_end() frees them."""

    def __init__(self, eng, ids, names):
        self.eng, self.ids, self.names = eng, list(ids), list(names)
        self.generation = eng._msa_generation
        self.data = None
        track = getattr(eng, "msa_track", None)          # stand-in engines of the CPU tests have no pool to replace
        if track is not None:
            track(self)

    def fetch(self):
        if self.data is None:
            if self.eng._msa_generation != self.generation:
                raise RuntimeError("the device pool of this alignment was replaced before its nodes were fetched")
            self.data = self.eng.msa_fetch(self.ids)
        return self.data


class _LeafWeights(collections.abc.Sequence):
    """The consensus weights of the leaves (multiple_alignment.py:184-188): np.full((len, 1), consensus_weight), made when read."""

    def __init__(self, lengths, consensus_weight):
        self._lengths, self._w = list(lengths), float(consensus_weight)

    def __len__(self):
        return len(self._lengths)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        return np.full((self._lengths[i], 1), self._w, dtype=np.float64)


class _LazyNodeList(collections.abc.Sequence):
    """final_sequences / final_consensus_weights (multiple_alignment.py:251-252): the leaves followed by the intermediate nodes."""

    def __init__(self, head, nodes: _PoolNodes, make):
        self._head, self._nodes, self._make, self._tail = head if isinstance(head, _LeafWeights) else list(head), nodes, make, None

    def _materialise(self):
        if self._tail is None:
            self._tail = [self._make(q, rec) for q, rec in enumerate(self._nodes.fetch())]
        return self._tail

    def __len__(self):
        return len(self._head) + len(self._nodes.ids)

    def __reduce__(self):
        return list, (list(self._head) + self._materialise(),)        # pickles as the reference's plain list (nodes fetched now)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return (list(self._head) + self._materialise())[i]
        if i < 0:
            i += len(self)
        if i < 0 or i >= len(self):
            raise IndexError(i)
        return self._head[i] if i < len(self._head) else self._materialise()[i - len(self._head)]


@dataclass
class MultipleAlignment:
    """The part of the reference's MultipleAlignment that sits on the hot path (multiple_alignment.py:148-170)."""
    sequences: typing.List[typing.Any]
    tree: typing.Optional[np.ndarray] = None
    branch_lengths: typing.Optional[np.ndarray] = None
    alignment: typing.Optional[typing.Dict[str, np.ndarray]] = None
    precision: typing.Optional[int] = None
    last_status: typing.Optional[np.ndarray] = field(default=None, repr=False)
    final_consensus_weights: typing.Optional[list] = field(default=None, repr=False)
    final_alignments: typing.Optional[dict] = field(default=None, repr=False)
    final_sequences: typing.Optional[list] = field(default=None, repr=False)

    def _params(self, score_function_params) -> _engine.Params:
        p = dict(DEFAULT_SCORE_PARAMS)
        # the reference's Protein.score_function defaults (gamma_tensor=0.03) apply when the caller passes nothing
        # (:321-322); align_from_structure_files always passes 7.0 / 0.03 (:490-492)
        if score_function_params is None:
            p.update(gamma_tensor=0.03, gamma_coords=0.03)
        else:
            unknown = set(score_function_params) - {"flexible", "gamma_tensor", "gamma_coords", "verbose"}
            if unknown:
                raise TypeError(f"score_function() got unexpected keyword arguments {sorted(unknown)}")
            if "gamma_tensor" not in score_function_params:
                p["gamma_tensor"] = 0.03
            if "gamma_coords" not in score_function_params:
                p["gamma_coords"] = 0.03
            p.update(score_function_params)
        prec = self.precision if self.precision is not None else _precision_from_env()
        # flexible=True (:323-326): the score matrix is the tensor Gaussian alone -> smith_waterman_score of it, gamma_coords unused
        return _engine.Engine.params(p["gamma_tensor"], p["gamma_coords"], prec, flexible=bool(p.get("flexible", False)))

    def make_pairwise_matrix(self, score_function_params=None) -> np.ndarray:
        """float64 [N,N] similarity, symmetric, zero diagonal (the caller turns it into max - S, :501)."""
        eng = get_engine()
        if not _on_fused_path(self.sequences):
            return self._pairwise_matrix_generic(score_function_params or {})
        prm = self._params(score_function_params)
        packed = _pack_staged(self.sequences, need_coordinates=not prm.flags & _engine.FLAG_FLEXIBLE)
        n = len(self.sequences)
        lens = np.diff(packed[2]).astype(np.float64)
        cells = 0.5 * (lens.sum() ** 2 - (lens ** 2).sum())
        multi = get_multi_engine() if cells >= MULTI_GPU_MIN_CELLS else None
        if multi is not None:
            # all GPUs of the box: shards by cost, one all-gather, bitwise the one-GPU matrix (tests/test_gpu_multi.py)
            multi.set_chains(*packed)
            return multi.pairwise_all(prm)
        eng.set_chains(*packed)
        if n < 2:
            return np.zeros((n, n))
        return eng.pairwise_all(prm)

    def _pairwise_matrix_generic(self, params, batch_bytes: int = 256 << 20) -> np.ndarray:
        """Sequences of any other SequenceBase type (:109-127): their own score_function makes each matrix on the host, the
        Smith-Waterman scores (dynamic_time_warping.py:204-222) are computed on the device in batches of matrices."""
        eng = get_engine()
        n = len(self.sequences)
        out = np.zeros((n, n))
        todo, mats, size = [], [], 0

        def flush():
            nonlocal todo, mats, size
            for (i, j), res in zip(todo, eng.sw_align_batch(mats, 0.0, want_paths=False)):
                out[i, j] = out[j, i] = res[2]
            todo, mats, size = [], [], 0

        for i in range(n - 1):
            for j in range(i + 1, n):
                m = np.ascontiguousarray(self.sequences[i].score_function(self.sequences[j], **params), dtype=np.float64)
                if m.shape != (len(self.sequences[i]), len(self.sequences[j])):
                    raise ValueError(f"score_function of {self.sequences[i].name} returned shape {m.shape}")
                todo.append((i, j)); mats.append(m); size += m.nbytes
                if size >= batch_bytes:
                    flush()
        if todo:
            flush()
        return out

    def _progressive_align_generic(self, tree, gap_open_penalty, gap_extend_penalty, consensus_weight, gamma_weight, sparams, mparams):
        """progressive_align (:172-253) for sequences without shape tensors: score_function / mean_function are the sequences'
        own; the weight Gaussian (:206-210), the affine DTW (:211-214) and get_mean_weights (:217) run on the device."""
        eng = get_engine()
        nodes = list(self.sequences)
        weights = [np.full((len(s), 1), consensus_weight, dtype=np.float64) for s in nodes]
        members = {s.name: {s.name: np.arange(len(s))} for s in nodes}

        def join(a, b, label):
            sa, sb = nodes[a], nodes[b]
            ka, kb = len(members[sa.name]), len(members[sb.name])
            S = np.array(sa.score_function(sb, **sparams), dtype=np.float64)
            S += eng.score_matrix(weights[a] * (kb / (2 * (ka + kb))), None, weights[b] * (ka / (2 * (ka + kb))), None,
                                  gamma_weight, 0.0, flexible=True)[0]
            al_a, al_b, _ = eng.dtw_align_batch([S], gap_open_penalty, gap_extend_penalty)[0]
            al_a, al_b = al_a.astype(np.int64), al_b.astype(np.int64)
            merged = {}
            for side, al in ((sa.name, al_a), (sb.name, al_b)):
                ext = {k: np.where(al >= 0, np.asarray(v)[np.maximum(al, 0)], -1) for k, v in members[side].items()}
                members[side] = ext
                merged.update(ext)
            name = f"int-{label}"
            members[name] = merged
            nodes.append(sa.mean_function(sb, al_a, al_b, name, **mparams))
            weights.append(eng.mean_weights(weights[a], weights[b], al_a, al_b))

        tree = np.asarray(tree)
        for x in range(0, tree.shape[0] - 1, 2):
            assert int(tree[x + 1, 1]) == int(tree[x, 1])
            join(int(tree[x, 0]), int(tree[x + 1, 0]), int(tree[x, 1]))
        a, b = int(tree[-1, 0]), int(tree[-1, 1])
        join(a, b, "final")
        self.final_consensus_weights, self.final_alignments, self.final_sequences = weights, members, nodes
        return {**members[nodes[a].name], **members[nodes[b].name]}

    # ------------------------------------------------------------------ guide tree + progressive alignment
    def progressive_align(self, tree, gap_open_penalty, gap_extend_penalty, consensus_weight, gamma_weight,
                          score_function_params=None, mean_function_params=None) -> typing.Dict[str, np.ndarray]:
        """MultipleAlignment.progressive_align (multiple_alignment.py:172-253): same bookkeeping and the same results; every node's
        score matrix, affine DTW and intermediate node are computed on the device.  A node depends only on its two children, so all
        nodes of one dependency level of the guide tree go to the device in ONE call instead of the reference's one-node-at-a-time
        loop, and the sequences (leaves and intermediate nodes) stay in a pool on the device (crt_msa_*): per level only the
        alignments come back; final_sequences / final_consensus_weights are fetched when they are first read.
        CARETTA_B200_MSA_POOL=0: crt_progressive_level with host arrays per level; CARETTA_B200_NODE_BATCH=0: one
        crt_progressive_node call per node, in tree order."""
        p = dict(score_function_params or {})
        if not _on_fused_path(self.sequences):
            return self._progressive_align_generic(tree, gap_open_penalty, gap_extend_penalty, consensus_weight, gamma_weight, p,
                                                   dict(mean_function_params or {}))
        flex_score, flex_mean = bool(p.get("flexible", False)), bool((mean_function_params or {}).get("flexible", False))
        gt, gc = p.get("gamma_tensor", 0.03), p.get("gamma_coords", 0.03)          # Protein.score_function defaults, :321-322
        if flex_mean and not flex_score and len(self.sequences) > 2:
            # the reference makes coordinate-less nodes (:359-360) and then fails scoring them with flexible=False (:337-349)
            raise ValueError("mean_function_params flexible=True needs score_function_params flexible=True: a flexible node has no "
                             "coordinates to score on")
        if flex_score:                             # tensor-only score matrices; the C ABI's sentinels for gamma_coords
            gc = _engine.GC_FLEXIBLE if flex_mean else _engine.GC_FLEXIBLE_SCORE
        need_xyz = not (flex_score and flex_mean)

        def make_node(name, tensors, coordinates):
            return Protein(name, tensors) if flex_mean else Protein(name, tensors, coordinates)
        eng = get_engine()
        n_leaves = len(self.sequences)
        tree = np.asarray(tree)
        # (node_1, node_2, name) in the reference's creation order; the node made at step q gets index n_leaves + q (:236-247)
        steps = []
        for x in range(0, tree.shape[0] - 1, 2):
            node_1, node_2, node_int = int(tree[x, 0]), int(tree[x + 1, 0]), int(tree[x, 1])
            assert int(tree[x + 1, 1]) == node_int
            steps.append((node_1, node_2, f"int-{node_int}"))
        steps.append((int(tree[-1, 0]), int(tree[-1, 1]), "int-final"))
        n_total = n_leaves + len(steps)
        final_sequences = [s for s in self.sequences] + [None] * len(steps)
        leaf_lengths = [len(s) for s in self.sequences]
        use_pool = os.environ.get("CARETTA_B200_NODE_BATCH", "1") != "0" and os.environ.get("CARETTA_B200_MSA_POOL", "1") != "0" \
            and hasattr(eng, "msa_level")
        # (:184-188) with the sequences in the device pool the leaves' weights are only made when somebody reads them
        final_consensus_weights = None if use_pool else \
            [np.full((n, 1), consensus_weight, dtype=np.float64) for n in leaf_lengths] + [None] * len(steps)
        # Bookkeeping.  The reference re-indexes the index arrays of EVERY member of both children at every node (:219-226), which
        # is O(N x depth x length) in total.  Here a node only keeps its two alignments; index arrays are composed top-down when
        # they are needed: the map (columns of a frame -> columns of node j) goes down the tree with one gather per edge, and at
        # a leaf the map IS the index array.  The final alignment costs O(nodes x length); any node's final_alignments entry
        # costs O(its subtree) when it is read.
        level = [0] * n_leaves + [0] * len(steps)
        count = [1] * n_leaves + [0] * len(steps)                     # sequences under a node = len(final_alignments[name]) (:200-203)
        for q, (a, b, _) in enumerate(steps):
            if not (0 <= a < n_leaves + q and 0 <= b < n_leaves + q):
                raise IndexError("tree refers to a node that does not exist yet")
            level[n_leaves + q] = 1 + max(level[a], level[b])
            count[n_leaves + q] = count[a] + count[b]
        statuses = np.zeros(len(steps), np.int32)
        node_len = leaf_lengths + [0] * len(steps)
        down = [None] * n_total                   # node -> (aln_1, aln_2) with a -1 sentinel appended: index -1 (gap) picks -1
        parent_side = [None] * n_total            # child -> its side of the parent's alignment (columns of the parent -> columns of the child)
        with_sentinel = False

        def finish(q, res):
            a, b, name_int = steps[q]
            i = n_leaves + q
            statuses[q] = res[-1]
            if with_sentinel:                                       # msa_level_ext: views that already end in the -1 sentinel
                ext = (res[0], res[1])
            else:
                ext = []
                for al in (res[0], res[1]):
                    e = np.empty(len(al) + 1, np.int32)
                    e[:-1] = al
                    e[-1] = -1
                    ext.append(e)
            down[i] = (ext[0], ext[1])
            parent_side[a], parent_side[b] = ext[0][:-1], ext[1][:-1]
            node_len[i] = len(ext[0]) - 1
            if len(res) > 4:                                        # host path: the node itself comes back with the alignment
                final_sequences[i] = make_node(name_int, res[2], res[3])
                final_consensus_weights[i] = res[4]

        def multipliers(q):
            a, b, _ = steps[q]
            l1, l2 = count[a], count[b]
            return (l2 / (2 * (l1 + l2)), l1 / (2 * (l1 + l2)))         # multiplier_n1, multiplier_n2 (:200-203)

        def node_inputs(q):
            a, b, _ = steps[q]
            s1, s2 = final_sequences[a], final_sequences[b]
            xyz = [s.coordinates if need_xyz or getattr(s, "coordinates", None) is not None else np.zeros((len(s), 3)) for s in (s1, s2)]
            return ((s1.tensors, xyz[0], final_consensus_weights[a]), (s2.tensors, xyz[1], final_consensus_weights[b])), \
                multipliers(q)

        levels = [[] for _ in range(max(level) if steps else 0)]
        for q in range(len(steps)):
            levels[level[n_leaves + q] - 1].append(q)
        if use_pool:
            # the sequences stay on the device: leaves = pool ids 0..N-1, every level appends its nodes; only alignments come back
            eng.set_chains(*_pack_staged(self.sequences, need_coordinates=need_xyz))
            eng.msa_begin(consensus_weight)
            pool_id = list(range(n_leaves)) + [None] * len(steps)
            level_call = getattr(eng, "msa_level_ext", None)          # alignments come back with the sentinel in place: no copies
            with_sentinel = level_call is not None
            if level_call is None:
                level_call = eng.msa_level
            for qs in levels:
                first, results = level_call([pool_id[steps[q][0]] for q in qs], [pool_id[steps[q][1]] for q in qs],
                                            [multipliers(q) for q in qs], gt, gc, gamma_weight, gap_open_penalty, gap_extend_penalty)
                for k, (q, res) in enumerate(zip(qs, results)):
                    pool_id[n_leaves + q] = first + k
                    finish(q, res)
            nodes = _PoolNodes(eng, pool_id[n_leaves:], [st[2] for st in steps])
            final_sequences = _LazyNodeList(final_sequences[:n_leaves], nodes, lambda q, rec: make_node(steps[q][2], rec[0], rec[1]))
            final_consensus_weights = _LazyNodeList(_LeafWeights(leaf_lengths, consensus_weight), nodes, lambda q, rec: rec[2])
            if os.environ.get("CARETTA_B200_FETCH_NODES", "0") != "0":
                nodes.fetch()
        elif os.environ.get("CARETTA_B200_NODE_BATCH", "1") != "0":
            for qs in levels:
                inputs = [node_inputs(q) for q in qs]
                results = eng.progressive_level([x[0] for x in inputs], [x[1] for x in inputs], gt, gc, gamma_weight,
                                                gap_open_penalty, gap_extend_penalty)
                for q, res in zip(qs, results):
                    finish(q, res)
        else:
            for q in range(len(steps)):
                (c1, c2), (m1, m2) = node_inputs(q)
                finish(q, eng.progressive_node(*c1, *c2, m1, m2, gt, gc, gamma_weight, gap_open_penalty, gap_extend_penalty))

        node_names = [s.name for s in self.sequences] + [st[2] for st in steps]

        def leaf_maps(i, frame_map):
            """[(leaf, int64 index array)] of the sequences under node i, first child's first (the dict merge order of :229-232);
            frame_map: columns of the frame -> columns of node i (-1 = gap)."""
            out, stack = [], [(i, frame_map)]
            while stack:
                j, mp = stack.pop()
                if j < n_leaves:
                    out.append((j, mp.astype(np.int64)))            # a leaf's columns are its residue indices
                    continue
                a, b = steps[j - n_leaves][0], steps[j - n_leaves][1]
                stack.append((b, np.take(down[j][1], mp)))
                stack.append((a, np.take(down[j][0], mp)))
            return out

        def node_alignments(i):
            # every node's entry is re-written in its parent's frame when the parent is made (:219-226); the last node keeps its own
            frame = parent_side[i] if parent_side[i] is not None else np.arange(node_len[i], dtype=np.int32)
            return {node_names[leaf]: arr for leaf, arr in leaf_maps(i, frame)}

        final_alignments = _NodeAlignments(node_names, node_alignments)
        last = n_total - 1
        if use_pool and steps and hasattr(eng, "msa_compose") and os.environ.get("CARETTA_B200_MSA_COMPOSE", "1") != "0":
            # the same top-down composition inside the library (crt_msa_compose): rows of one [N, A] int64 array, dictionary order
            ids, rows = eng.msa_compose(pool_id[last])
            alignment = {node_names[leaf]: rows[r] for r, leaf in enumerate(ids.tolist())}
        else:
            alignment = {node_names[leaf]: arr for leaf, arr in leaf_maps(last, np.arange(node_len[last], dtype=np.int32))}
        self.final_consensus_weights = final_consensus_weights
        self.final_alignments = final_alignments
        self.final_sequences = final_sequences
        self.last_status = statuses
        return alignment

    def multiple_align(self, pairwise_distance_matrix, gap_open_penalty, gap_extend_penalty, consensus_weight, gamma_weight,
                       score_function_params=None, mean_function_params=None) -> typing.Dict[str, np.ndarray]:
        """MultipleAlignment.multiple_align (multiple_alignment.py:255-285): neighbor joining + progressive alignment."""
        from . import neighbor_joining as _nj
        if len(self.sequences) == 2 and not _on_fused_path(self.sequences):
            s1, s2 = self.sequences
            S = np.ascontiguousarray(s1.score_function(s2, **(score_function_params or {})), dtype=np.float64)
            aln_1, aln_2, _ = get_engine().dtw_align_batch([S], gap_open_penalty, gap_extend_penalty)[0]
            self.alignment = {s1.name: aln_1.astype(np.int64), s2.name: aln_2.astype(np.int64)}
            return self.alignment
        if len(self.sequences) == 2:
            p = dict(score_function_params or {})
            s1, s2 = self.sequences
            # two structures: dtw_align on the plain score matrix (:263-275); gamma_weight < 0 switches the weight term off
            flex = bool(p.get("flexible", False))                  # tensor-only score matrix (:323-326); coordinates are not read
            xyz = [np.zeros((len(s), 3)) if flex and getattr(s, "coordinates", None) is None else s.coordinates for s in (s1, s2)]
            aln_1, aln_2, *_ = get_engine().progressive_node(
                s1.tensors, xyz[0], np.zeros(len(s1)), s2.tensors, xyz[1], np.zeros(len(s2)), 0.0, 0.0,
                p.get("gamma_tensor", 0.03), _engine.GC_FLEXIBLE if flex else p.get("gamma_coords", 0.03), -1.0,
                gap_open_penalty, gap_extend_penalty)
            self.alignment = {s1.name: aln_1, s2.name: aln_2}
            return self.alignment
        self.tree, self.branch_lengths = _nj.neighbor_joining(pairwise_distance_matrix)
        self.alignment = self.progressive_align(self.tree, gap_open_penalty, gap_extend_penalty, consensus_weight, gamma_weight,
                                                score_function_params, mean_function_params)
        return self.alignment

    # ------------------------------------------------------------------ text writers
    def _fasta_bytes(self, alignment=None) -> bytes:
        if alignment is None:
            alignment = self.alignment
        names = [p.name for p in self.sequences]
        aln = np.array([np.asarray(alignment[n], dtype=np.int64) for n in names])
        seqs = [str(p) for p in self.sequences]
        for q in seqs:
            if not q.isascii():
                raise ValueError("sequences must be ASCII (one byte per residue)")
        return get_engine().format_fasta(names, seqs, aln)

    def to_sequence_alignment(self, alignment=None) -> typing.Dict[str, str]:
        """multiple_alignment.py:287-297: {name: aligned amino-acid string with '-' for gaps}; the gather runs on the device."""
        text = self._fasta_bytes(alignment)
        out, pos = {}, 0
        alen = len(next(iter((alignment or self.alignment).values()))) if len(self.sequences) else 0
        for p in self.sequences:
            nl = len(p.name.encode("utf-8"))
            start = pos + 1 + nl + 1
            out[p.name] = text[start:start + alen].decode("ascii")
            pos = start + alen + 1
        return out

    def write_alignment(self, fasta_file, alignment=None) -> None:
        """multiple_alignment.py:299-309: same bytes as the reference's file."""
        text = self._fasta_bytes(alignment)
        with open(fasta_file, "wb") as f:
            f.write(text)

    def make_pairwise_matrices(self, score_function_params=None):
        """Engine by-product: (score, rmsd, tm) over the stage-1 matched residues, each float64 [N,N]."""
        eng = get_engine()
        eng.set_chains(*_pack_staged(self.sequences))
        return eng.pairwise_all(self._params(score_function_params), want_rmsd_tm=True)


class StructureMultiple(MultipleAlignment):
    """Name kept for the legacy API / the CLI help text (bin/caretta-cli:84)."""

    @classmethod
    def from_arrays(cls, names, tensors_list, coords_list, sequences=None):
        seqs = sequences or ["" for _ in names]
        return cls([Protein(n, np.asarray(t, np.float64), np.asarray(c, np.float64), s)
                    for n, t, c, s in zip(names, tensors_list, coords_list, seqs)])

    @classmethod
    def from_chains(cls, chains):
        return cls([Protein(f"s{p}", *chains.chain(p), "") for p in range(chains.n)])

    def make_pairwise_score_matrix(self, **score_function_params):
        return self.make_pairwise_matrix(score_function_params or dict(DEFAULT_SCORE_PARAMS))


# ---------------------------------------------------------------------------------------------------------------
# The DPs in isolation, with the reference's signatures (dynamic_time_warping.py:147-278).  seq1/seq2 must be the
# identity index arrays the reference always passes on this path (np.arange(n), np.arange(m)).
# ---------------------------------------------------------------------------------------------------------------
def _check_arange(seq, n, what):
    seq = np.asarray(seq)
    if len(seq) != n or (n and not np.array_equal(seq, np.arange(n))):
        raise NotImplementedError(f"{what}: only the identity indexing np.arange({n}) used on caretta's pair path is accelerated")


def dtw_align(seq1, seq2, score_matrix, gap_open_penalty: float = 0.0, gap_extend_penalty: float = 0.0):
    score_matrix = np.asarray(score_matrix, dtype=np.float64)
    _check_arange(seq1, score_matrix.shape[0], "dtw_align")
    _check_arange(seq2, score_matrix.shape[1], "dtw_align")
    a1, a2, sc = get_engine().dtw_align_batch([score_matrix], gap_open_penalty, gap_extend_penalty)[0]
    return a1, a2, sc


def dtw_align_score(seq1, seq2, score_matrix, gap_open_penalty: float = 0.0, gap_extend_penalty: float = 0.0):
    return dtw_align(seq1, seq2, score_matrix, gap_open_penalty, gap_extend_penalty)[2]


def smith_waterman(seq1, seq2, score_matrix, gap: float = 0.0):
    score_matrix = np.asarray(score_matrix, dtype=np.float64)
    _check_arange(seq1, score_matrix.shape[0], "smith_waterman")
    _check_arange(seq2, score_matrix.shape[1], "smith_waterman")
    a1, a2, sc, st = get_engine().sw_align_batch([score_matrix], gap)[0]
    if st & _engine.ST_NO_POSITIVE:
        # the reference fails here too (max_pos is None -> TypeError inside numba, dynamic_time_warping.py:250)
        raise TypeError("smith_waterman: no cell of the score matrix is positive")
    return a1, a2, sc


def smith_waterman_score(seq1, seq2, matrix, gap: float = 0.0):
    matrix = np.asarray(matrix, dtype=np.float64)
    _check_arange(seq1, matrix.shape[0], "smith_waterman_score")
    _check_arange(seq2, matrix.shape[1], "smith_waterman_score")
    return get_engine().sw_align_batch([matrix], gap, want_paths=False)[0][2]


def make_rmsd_coverage_tm_matrix(alignment, proteins, superpose_first: bool = True):
    """multiple_alignment.py:1000-1055.  alignment is {name: int64[A]} with -1 gaps, proteins the matching list of Protein-like
    objects.  superpose_first=True superposes all structures with superpose() first (and, like the reference, leaves the
    proteins with their new coordinates), then measures every pair in that common frame."""
    if superpose_first:
        proteins = superpose(alignment, proteins)
    names = [p.name for p in proteins]
    aln = np.array([np.asarray(alignment[n], dtype=np.int64) for n in names])
    eng = get_engine()
    eng.set_coords(*pack_coordinates(proteins))
    r, c, t, bad = eng.rmsd_cov_tm(aln, superpose=not superpose_first)
    assert bad == 0, "a pair has fewer than 3 common positions (the reference asserts here, :1034)"
    return r, c, t


# ---------------------------------------------------------------------------------------------------------------
# Consumers of the alignment: coverage matrix, reference selection, superposition.
# ---------------------------------------------------------------------------------------------------------------
def _aln_array(alignment, names):
    return np.array([np.asarray(alignment[n], dtype=np.int64) for n in names])


def _check_gap(gap):
    if gap != -1:
        raise NotImplementedError("only the reference's gap marker -1 is supported")


def make_coverage_gap_distance_matrix(alignment_array):
    """multiple_alignment.py:45-56: (distance float64 [N,N], matrix_aligning int32 [N,N])."""
    return get_engine().coverage_gap_matrix(np.asarray(alignment_array))


def get_reference_structures(alignment, minimum_coverage=50, gap=-1):
    """multiple_alignment.py:740-784: a set of reference structures such that every structure shares at least
    minimum_coverage % of its residues with the reference it is assigned to.  The O(N^2 A) counting runs on the device
    (crt_coverage_gap_matrix); the greedy choice over the two N x N matrices is host logic with the reference's rules:
    the first reference minimises the median gap fraction, every further one is the already assigned structure with the
    smallest median gap fraction towards the structures still waiting (plain minimum when only one waits).
    Returns (first reference name, {reference name: [member names]}, [names that align to nobody])."""
    _check_gap(gap)
    names = list(alignment)
    aln = _aln_array(alignment, names)
    gap_fraction, shared = make_coverage_gap_distance_matrix(aln)
    need = minimum_coverage * (aln != gap).sum(axis=1) / 100            # residues a structure must share with its reference

    def split(ref, who):
        ok = shared[who, ref] >= need[who]
        return who[ok], who[~ok]

    groups = {}                                                         # reference index -> member indices, insertion order
    first = int(np.argmin(np.median(gap_fraction, axis=0)))
    taken, waiting = split(first, np.arange(len(names)))
    groups[first] = [int(i) for i in taken]
    assigned = list(groups[first])
    stranded = []
    while waiting.size:
        towards = gap_fraction[np.ix_(waiting, assigned)]
        pick = np.argmin(np.median(towards, axis=0)) if waiting.size > 1 else np.argmin(towards)
        ref = assigned[int(pick)]
        taken, rest = split(ref, waiting)
        if taken.size == 0:                                             # nobody can take the rest: try them one by one below
            stranded = [int(i) for i in waiting]
            break
        groups[ref] = [int(i) for i in taken]
        assigned += groups[ref]
        waiting = rest
    alone = []
    for i in stranded:
        home = next((j for j in assigned if shared[i, j] >= need[i]), None)
        if home is None:
            alone.append(names[i])
        else:
            groups[home].append(i)
    return names[first], {names[r]: [names[i] for i in members] for r, members in groups.items()}, alone


def _superpose_on_device(alignment, proteins, mode, reference_name=None, core_indices=None):
    names = [p.name for p in proteins]
    eng = get_engine()
    eng.set_coords(*pack_coordinates(proteins))
    ref = -1 if reference_name is None else names.index(reference_name)
    res = eng.superpose(_aln_array(alignment, names), mode, ref, core_indices)
    off = eng._offsets
    if res["mode"] == _engine.SUP_REFERENCE:
        assert int(res["ncommon"].min()) > 3, "a structure has <= 3 positions in common with the reference (the reference asserts, :918)"
    for p, prot in enumerate(proteins):
        prot.coordinates = res["coords"][off[p]:off[p + 1]].copy()
    return proteins, res


def superpose(alignment, proteins, gap=-1):
    """multiple_alignment.py:854-867: core superposition when at least half of the columns are gap-free, else onto the
    reference structure (the protein with the most aligned residues)."""
    _check_gap(gap)
    # the reference's own rule, on the host: the reference structure is the FIRST key of `alignment` (its order is
    # the guide-tree traversal order, not the order of `proteins`) among those with the most aligned residues -- sorted() is
    # stable, also with reverse=True --, and the core columns are gap-free over EVERY entry of the alignment
    keys = list(alignment.keys())
    rows = np.stack([np.asarray(alignment[k]) for k in keys])
    reference_name = sorted(keys, key=lambda k: int(np.count_nonzero(np.asarray(alignment[k]) != -1)), reverse=True)[0]
    core_indices = np.nonzero(np.all(rows != -1, axis=0))[0]
    print("Core indices", len(core_indices))                               # the reference prints this, :906
    if len(core_indices) < rows.shape[1] // 2:
        return superpose_reference(alignment, proteins, reference_name)
    return superpose_core(alignment, proteins, reference_name, core_indices)


def superpose_core(alignment, proteins, reference_name, core_indices: np.ndarray = None, gap=-1):
    """multiple_alignment.py:869-905."""
    _check_gap(gap)
    return _superpose_on_device(alignment, proteins, _engine.SUP_CORE, reference_name, core_indices)[0]


def superpose_reference(alignment, proteins, reference_name):
    """multiple_alignment.py:908-927."""
    return _superpose_on_device(alignment, proteins, _engine.SUP_REFERENCE, reference_name)[0]


def superpose_references(alignment, proteins, minimum_coverage=50):
    """multiple_alignment.py:930-950: every group of get_reference_structures is superposed onto its reference, group after
    group (a later reference has already been moved by an earlier group); one device call, one launch per dependency level."""
    names = [p.name for p in proteins]
    index = {n: q for q, n in enumerate(names)}
    first_reference_structure, reference_structures, no_aligning = get_reference_structures(alignment, minimum_coverage)
    ref, mem, batch_off = [], [], [0]
    for reference_name, members in reference_structures.items():
        r = index[reference_name]
        # members in the reference's loop order; the reference itself (if listed) is replaced on its turn, so the members
        # before it see the old coordinates and the ones after it the new ones: up to three dependent batches
        cut = [q for q, n in enumerate(members) if index[n] == r]
        parts = [members] if not cut else [members[:cut[0]], members[cut[0]:cut[0] + 1], members[cut[0] + 1:]]
        for part in parts:
            if part:
                ref += [r] * len(part)
                mem += [index[n] for n in part]
                batch_off.append(len(ref))
    eng = get_engine()
    eng.set_coords(*pack_coordinates(proteins))
    res = eng.superpose_pairs(_aln_array(alignment, names), ref, mem, batch_off)
    assert len(ref) == 0 or int(res["ncommon"].min()) > 3, "a structure has <= 3 positions in common with its reference (:941)"
    off = eng._offsets
    for p, prot in enumerate(proteins):
        prot.coordinates = res["coords"][off[p]:off[p + 1]].copy()
    return proteins


def make_count_matrix(residues_list, alphabet_size: int) -> np.ndarray:
    """multiple_alignment.py:128-134 (fast-mode guide matrix, :503-509): shapemer counts per protein."""
    return get_engine().count_matrix(residues_list, alphabet_size)


def braycurtis(counts_1, counts_2) -> np.ndarray:
    """multiple_alignment.py:137-145: Bray-Curtis distance between every row of counts_1 and every row of counts_2."""
    return get_engine().braycurtis(counts_1, counts_2)


def write_distance_matrix(names, distance_matrix, filename) -> None:
    """helper.write_distance_matrix (helper.py:183-203): Clustal-style text, '%.4f' values formatted on the device, same bytes."""
    text = get_engine().format_matrix([str(n) for n in names], np.asarray(distance_matrix, dtype=np.float64), view=True)
    with open(filename, "wb") as f:
        f.write(text)                       # straight from the engine's page-locked staging buffer


def read_distance_matrix(filename):
    """helper.read_distance_matrix (helper.py:205-229): the inverse of write_distance_matrix -- (names, float64 [N,N]).  File
    parsing (host I/O, not part of the compute path): first line N, then 'name v v ...' rows; a name is cut at its first '/'."""
    with open(filename, "rb") as f:
        lines = f.read().split(b"\n")
    n = int(lines[0].strip())
    rows = [ln.split() for ln in lines[1:] if ln.strip()]
    names = [r[0].decode("utf-8").strip().split("/")[0].strip() for r in rows]
    assert len(names) == n
    matrix = np.array([[float(x) for x in r[1:n + 1]] for r in rows], dtype=np.float64)
    return names, matrix


def alignment_to_numpy(alignment) -> typing.Dict[str, np.ndarray]:
    """multiple_alignment.py:30-42: {name: aligned string with '-'} -> {name: residue index per column, -1 = gap}."""
    out = {}
    for name, text in alignment.items():
        residue = np.frombuffer(text.encode("utf-8") if isinstance(text, str) else bytes(text), dtype=np.uint8) != ord("-")
        out[name] = np.where(residue, np.cumsum(residue) - 1, -1).astype(np.int64)
    return out


def get_gaussian_score(coord_1, coord_2, gamma=0.03):
    """score_functions.get_gaussian_score (score_functions.py:6-11) for two points; as the score_function argument of
    make_score_matrix it selects the device kernel."""
    return make_score_matrix(np.asarray(coord_1, dtype=np.float64).reshape(1, -1), np.asarray(coord_2, dtype=np.float64).reshape(1, -1),
                             get_gaussian_score, gamma)[0, 0]


def make_score_matrix(coords_1, coords_2, score_function=get_gaussian_score, gamma=0.03, normalized=False) -> np.ndarray:
    """score_functions.make_score_matrix (score_functions.py:22-51): float64 [n,m] of exp(-gamma sum_k (x[a,k] - y[b,k])^2), in the
    reference's operation order, on the device (crt_score_matrix, flexible form).  Only the Gaussian score the reference uses is a
    device kernel; normalized=True (never used by the reference's callers) is not provided."""
    if score_function is not get_gaussian_score and getattr(score_function, "__name__", "") != "get_gaussian_score":
        raise NotImplementedError("make_score_matrix: only get_gaussian_score runs on the device")
    if normalized:
        raise NotImplementedError("make_score_matrix(normalized=True) is not used on caretta's path and is not provided")
    return get_engine().score_matrix(coords_1, None, coords_2, None, gamma, 0.0, flexible=True)[0]


@dataclass
class OutputFiles:
    """multiple_alignment.py:85-105 (the files this package can write; PDB / feature / class files stay the reference's)."""
    output_folder: typing.Any = None
    fasta_file: typing.Any = None
    matrix_folder: typing.Any = None
    class_file: typing.Any = None

    @classmethod
    def from_folder(cls, output_folder):
        from pathlib import Path
        output_folder = Path(output_folder)
        return cls(output_folder, fasta_file=output_folder / "result.fasta", matrix_folder=output_folder / "result_matrix",
                   class_file=output_folder / "result_class.pkl")


def align_from_proteins(proteins, gap_open_penalty: float = 1.0, gap_extend_penalty: float = 0.01, consensus_weight: bool = True,
                        output_folder=None, write_fasta: bool = False, write_matrix: bool = False, verbose: bool = False,
                        full: bool = True, shapemer_indices=None, alphabet_size: typing.Optional[int] = None, write_class: bool = False):
    """align_from_structure_files (multiple_alignment.py:394-596) from the point where the features exist (:488-491): the list of
    Protein(name, tensors, coordinates, sequence) that the reference builds from geometricus.  Same steps, same parameters and
    the same output files as the reference with full=True (the caretta-cli default): all-vs-all matrix -> max - S (:498-501) ->
    guide-tree distance text (:515-522) -> multiple_align (:524-533) -> result.fasta (:540-545) -> RMSD / coverage / TM matrices and
    their text files (:571-591).  full=False is --fast (:503-511) from given shapemer indices; write_class pickles the
    MultipleAlignment (:557-559).  Returns (MultipleAlignment, OutputFiles)."""
    from pathlib import Path
    output_files = OutputFiles() if output_folder is None else OutputFiles.from_folder(output_folder)
    if output_folder is not None:
        Path(output_files.output_folder).mkdir(exist_ok=True)
    msa_class = MultipleAlignment(list(proteins))
    score_function_params = dict(DEFAULT_SCORE_PARAMS)
    mean_function_params = dict(flexible=False)
    pairwise_distance_matrix = np.array([[0, 1], [1, 0]])
    if len(msa_class.sequences) > 2:
        if full:
            pairwise_distance_matrix = msa_class.make_pairwise_matrix(score_function_params=score_function_params)
            pairwise_distance_matrix = pairwise_distance_matrix.max() - pairwise_distance_matrix
        else:
            # --fast (:503-511): Bray-Curtis distances of the shapemer counts; the shapemer indices per protein (geometricus'
            # map_protein_to_shapemer_indices) and the number of shapemer keys are inputs here
            if shapemer_indices is None or alphabet_size is None:
                raise ValueError("full=False needs shapemer_indices (one index array per protein) and alphabet_size")
            count_matrix = make_count_matrix(list(shapemer_indices), int(alphabet_size))
            pairwise_distance_matrix = braycurtis(count_matrix, count_matrix)
    names = [s.name for s in msa_class.sequences]
    if write_matrix:
        Path(output_files.matrix_folder).mkdir(exist_ok=True)
        write_distance_matrix(names, pairwise_distance_matrix, Path(output_files.matrix_folder) / "distance_matrix_guide_tree.txt")
    msa_class.pairwise_distance_matrix = pairwise_distance_matrix
    alignment = msa_class.multiple_align(pairwise_distance_matrix, gap_open_penalty=gap_open_penalty, gap_extend_penalty=gap_extend_penalty,
                                         consensus_weight=float(consensus_weight), gamma_weight=1.0,
                                         score_function_params=score_function_params, mean_function_params=mean_function_params)
    if write_fasta:
        msa_class.write_alignment(output_files.fasta_file)
    if write_class:                                                         # :557-559
        import pickle
        with open(output_files.class_file, "wb") as f:
            pickle.dump(msa_class, f)
    if write_matrix:
        rmsd, coverage, tm = make_rmsd_coverage_tm_matrix(alignment, msa_class.sequences, superpose_first=False)
        for fname, M in (("rmsd.txt", rmsd), ("coverage.txt", coverage), ("tm.txt", tm)):
            write_distance_matrix(names, M, Path(output_files.matrix_folder) / fname)
    if verbose:
        print(f"aligned {len(names)} structures, alignment length {len(next(iter(alignment.values())))}")
    return msa_class, output_files


def install(reference_multiple_alignment_module) -> None:
    """Monkey-patches the reference module in place: its MultipleAlignment.make_pairwise_matrix (the all-vs-all
    loop, :158-170) is replaced by the GPU path.  Everything else (neighbor joining, progressive alignment,
    writers, CLI flags) stays the reference's code."""
    ref = reference_multiple_alignment_module

    def make_pairwise_matrix(self, score_function_params=None):
        return MultipleAlignment(self.sequences).make_pairwise_matrix(score_function_params)

    ref.MultipleAlignment.make_pairwise_matrix = make_pairwise_matrix

    if os.environ.get("CARETTA_B200_TREE", "1") != "0":
        # the guide tree and the progressive alignment as well (CARETTA_B200_TREE=0 keeps the reference's)
        from . import neighbor_joining as _nj
        if hasattr(ref, "nj"):
            _nj.install(ref.nj)

        def progressive_align(self, tree, gap_open_penalty, gap_extend_penalty, consensus_weight, gamma_weight,
                              score_function_params=None, mean_function_params=None):
            m = MultipleAlignment(self.sequences)
            aln = m.progressive_align(tree, gap_open_penalty, gap_extend_penalty, consensus_weight, gamma_weight,
                                      score_function_params, mean_function_params)
            self.final_consensus_weights, self.final_alignments = m.final_consensus_weights, m.final_alignments
            self.final_sequences = [ref.Protein(s.name, s.tensors, s.coordinates, getattr(s, "sequence", "")) for s in m.final_sequences]
            return aln

        ref.MultipleAlignment.progressive_align = progressive_align

    if os.environ.get("CARETTA_B200_CONSUMERS", "1") != "0":
        # what consumes the alignment (CARETTA_B200_CONSUMERS=0 keeps the reference's)
        for fn in (make_coverage_gap_distance_matrix, get_reference_structures, superpose, superpose_core, superpose_reference,
                   superpose_references, make_rmsd_coverage_tm_matrix, make_count_matrix, braycurtis):
            setattr(ref, fn.__name__, fn)

        def to_sequence_alignment(self, alignment=None):
            return MultipleAlignment(self.sequences, alignment=self.alignment).to_sequence_alignment(alignment)

        def write_alignment(self, fasta_file, alignment=None):
            return MultipleAlignment(self.sequences, alignment=self.alignment).write_alignment(fasta_file, alignment)

        ref.MultipleAlignment.to_sequence_alignment = to_sequence_alignment
        ref.MultipleAlignment.write_alignment = write_alignment
        if hasattr(ref, "helper"):
            ref.helper.write_distance_matrix = write_distance_matrix
