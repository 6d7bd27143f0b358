"""Host-side mirror of the reference's interface for the all-vs-all pair path (names, argument meaning and error
behaviour follow caretta/multiple_alignment.py of TurtleTools/caretta 0.2.0), backed by the CUDA engine through
the C ABI (caretta_b200.engine).  Nothing here computes the path on the CPU.

    reference                                                        here
    ---------------------------------------------------------------  --------------------------------------------
    Protein(name, tensors, coordinates, sequence)        :312-319    Protein (same fields, same __len__/__str__)
    MultipleAlignment(sequences)                         :148-156    MultipleAlignment
      .make_pairwise_matrix(score_function_params)       :158-170    same signature, float64 [N,N], symmetric, diag 0
    make_rmsd_coverage_tm_matrix(alignment, proteins,    :1000-1055  same signature (superpose_first=False only,
                                 superpose_first)                    the reference's own call site, :571-572)
    dtw.dtw_align / smith_waterman / smith_waterman_score            dtw_align / smith_waterman / smith_waterman_score
      (dynamic_time_warping.py:147-278)                              (batched: *_batch)
    StructureMultiple (name used by the CLI help and the legacy API) StructureMultiple facade

``install(reference_module)`` swaps the reference's own methods for these, so that an unmodified
``align_from_structure_files`` / ``caretta-cli`` run uses the GPU for the pair loop.
"""
from __future__ import annotations

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


@dataclass
class Protein:
    """Same fields as the reference's Protein (multiple_alignment.py:312-319)."""
    name: str
    tensors: np.ndarray
    coordinates: np.ndarray = None
    sequence: str = ""

    def __len__(self) -> int:
        return self.tensors.shape[0]

    def __str__(self) -> str:
        return self.sequence


def pack_sequences(sequences) -> typing.Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """List of Protein-like objects (anything with .tensors [L,d] and .coordinates [L,3]) -> packed chain set."""
    if len(sequences) == 0:
        raise ValueError("no sequences")
    d = int(np.asarray(sequences[0].tensors).shape[1])
    lens = []
    for s in sequences:
        t = np.asarray(s.tensors)
        c = np.asarray(s.coordinates)
        if t.ndim != 2 or t.shape[1] != d:
            raise ValueError(f"{getattr(s, 'name', '?')}: tensors must be [L,{d}]")
        if c.ndim != 2 or c.shape[1] != 3 or c.shape[0] != t.shape[0]:
            raise ValueError(f"{getattr(s, 'name', '?')}: coordinates must be [L,3] with the same L as tensors")
        lens.append(t.shape[0])
    offsets = np.zeros(len(sequences) + 1, np.int64)
    offsets[1:] = np.cumsum(lens)
    coords = np.concatenate([np.asarray(s.coordinates, dtype=np.float64) for s in sequences])
    tensors = np.concatenate([np.asarray(s.tensors, dtype=np.float64) for s in sequences])
    return coords, tensors, offsets


_shared_engine: typing.Optional[_engine.Engine] = None


def get_engine() -> _engine.Engine:
    """One engine (context) per process, on the current CUDA device (LOCAL_RANK when launched by torchrun)."""
    global _shared_engine
    if _shared_engine is None:
        dev = int(os.environ.get("CARETTA_B200_DEVICE", os.environ.get("LOCAL_RANK", "-1")))
        _shared_engine = _engine.Engine(dev)
    return _shared_engine


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
        if p.get("flexible", False):
            raise NotImplementedError("flexible=True (tensor-only scoring, multiple_alignment.py:323-326) is not on the "
                                      "all-vs-all path and is not accelerated")
        prec = self.precision if self.precision is not None else _precision_from_env()
        return _engine.Engine.params(p["gamma_tensor"], p["gamma_coords"], prec)

    def make_pairwise_matrix(self, score_function_params=None) -> np.ndarray:
        """float64 [N,N] similarity, symmetric, zero diagonal (the caller turns it into max - S, :501)."""
        eng = get_engine()
        eng.set_chains(*pack_sequences(self.sequences))
        prm = self._params(score_function_params)
        n = len(self.sequences)
        if n < 2:
            return np.zeros((n, n))
        return eng.pairwise_all(prm)

    # ------------------------------------------------------------------ guide tree + progressive alignment
    def progressive_align(self, tree, gap_open_penalty, gap_extend_penalty, consensus_weight, gamma_weight,
                          score_function_params=None, mean_function_params=None) -> typing.Dict[str, np.ndarray]:
        """MultipleAlignment.progressive_align (multiple_alignment.py:172-253): same bookkeeping, every node's score matrix,
        affine DTW and intermediate node computed on the device by crt_progressive_node."""
        p = dict(score_function_params or {})
        if p.get("flexible", False) or (mean_function_params or {}).get("flexible", False):
            raise NotImplementedError("flexible=True is not accelerated")
        gt, gc = p.get("gamma_tensor", 0.03), p.get("gamma_coords", 0.03)          # Protein.score_function defaults, :321-322
        eng = get_engine()
        final_sequences = [s for s in self.sequences]
        final_alignments = {s.name: {s.name: np.arange(len(s))} for s in final_sequences}
        final_consensus_weights = [np.full((len(s), 1), consensus_weight, dtype=np.float64) for s in final_sequences]
        statuses = []

        def make_intermediate_node(n1, n2, n_int):
            s1, s2 = final_sequences[n1], final_sequences[n2]
            l1, l2 = len(final_alignments[s1.name]), len(final_alignments[s2.name])
            multiplier_n1, multiplier_n2 = l2 / (2 * (l1 + l2)), l1 / (2 * (l1 + l2))
            aln_1, aln_2, tm, cm, wm, _, st = eng.progressive_node(
                s1.tensors, s1.coordinates, final_consensus_weights[n1], s2.tensors, s2.coordinates, final_consensus_weights[n2],
                multiplier_n1, multiplier_n2, gt, gc, gamma_weight, gap_open_penalty, gap_extend_penalty)
            statuses.append(st)
            name_int = f"int-{n_int}"
            final_alignments[s1.name] = {name: np.where(aln_1 != -1, seq[np.maximum(aln_1, 0)], -1)
                                         for name, seq in final_alignments[s1.name].items()}
            final_alignments[s2.name] = {name: np.where(aln_2 != -1, seq[np.maximum(aln_2, 0)], -1)
                                         for name, seq in final_alignments[s2.name].items()}
            final_alignments[name_int] = {**final_alignments[s1.name], **final_alignments[s2.name]}
            final_sequences.append(Protein(name_int, tm, cm))
            final_consensus_weights.append(wm)

        tree = np.asarray(tree)
        for x in range(0, tree.shape[0] - 1, 2):
            node_1, node_2, node_int = int(tree[x, 0]), int(tree[x + 1, 0]), int(tree[x, 1])
            assert int(tree[x + 1, 1]) == node_int
            make_intermediate_node(node_1, node_2, node_int)
        node_1, node_2 = int(tree[-1, 0]), int(tree[-1, 1])
        make_intermediate_node(node_1, node_2, "final")
        alignment = {**final_alignments[final_sequences[node_1].name], **final_alignments[final_sequences[node_2].name]}
        self.final_consensus_weights = final_consensus_weights
        self.final_alignments = final_alignments
        self.final_sequences = final_sequences
        self.last_status = np.array(statuses, np.int32)
        return alignment

    def multiple_align(self, pairwise_distance_matrix, gap_open_penalty, gap_extend_penalty, consensus_weight, gamma_weight,
                       score_function_params=None, mean_function_params=None) -> typing.Dict[str, np.ndarray]:
        """MultipleAlignment.multiple_align (multiple_alignment.py:255-285): neighbor joining + progressive alignment."""
        from . import neighbor_joining as _nj
        if len(self.sequences) == 2:
            p = dict(score_function_params or {})
            s1, s2 = self.sequences
            # two structures: dtw_align on the plain score matrix (:263-275); gamma_weight < 0 switches the weight term off
            aln_1, aln_2, *_ = get_engine().progressive_node(
                s1.tensors, s1.coordinates, np.zeros(len(s1)), s2.tensors, s2.coordinates, np.zeros(len(s2)), 0.0, 0.0,
                p.get("gamma_tensor", 0.03), p.get("gamma_coords", 0.03), -1.0, gap_open_penalty, gap_extend_penalty)
            self.alignment = {s1.name: aln_1, s2.name: aln_2}
            return self.alignment
        self.tree, self.branch_lengths = _nj.neighbor_joining(pairwise_distance_matrix)
        self.alignment = self.progressive_align(self.tree, gap_open_penalty, gap_extend_penalty, consensus_weight, gamma_weight,
                                                score_function_params, mean_function_params)
        return self.alignment

    def make_pairwise_matrices(self, score_function_params=None):
        """Engine by-product: (score, rmsd, tm) over the stage-1 matched residues, each float64 [N,N]."""
        eng = get_engine()
        eng.set_chains(*pack_sequences(self.sequences))
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
    """multiple_alignment.py:1000-1055.  Only superpose_first=False (the reference's own call site, :571-572) is
    accelerated; alignment is {name: int64[A]} with -1 gaps, proteins the matching list of Protein-like objects."""
    if superpose_first:
        raise NotImplementedError("superpose_first=True goes through superpose() (reference/core selection, "
                                  "multiple_alignment.py:596-852), which is outside the accelerated path")
    names = [p.name for p in proteins]
    aln = np.array([np.asarray(alignment[n], dtype=np.int64) for n in names])
    eng = get_engine()
    eng.set_chains(*pack_sequences(proteins))
    r, c, t, bad = eng.rmsd_cov_tm(aln)
    assert bad == 0, "a pair has fewer than 3 common positions (the reference asserts here, :1034)"
    return r, c, t


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
