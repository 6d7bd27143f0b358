/*
 * caretta_b200.h -- C ABI of the B200-native engine for caretta's all-vs-all pair path.
 *
 * The reference (TurtleTools/caretta 0.2.0, pure Python + numba) has no FFI layer; these entry points are what a
 * ctypes binding placed behind the reference's own Python seams binds (INTEGRATION.md shows the stub):
 *
 *   crt_set_chains            <- the List[Protein] held by MultipleAlignment.sequences
 *                                (caretta/multiple_alignment.py:148-156, Protein :312-319: tensors f64[L,d],
 *                                coordinates f64[L,3]), packed residue-major.
 *   crt_pairwise_all          <- MultipleAlignment.make_pairwise_matrix        (multiple_alignment.py:158-170)
 *                                = for i<j: smith_waterman_score(Protein.score_function(...))   (:321-349, :164-169)
 *   crt_pairwise_list         <- the same per-pair recipe on an explicit pair list (parity tests, sampled baselines);
 *                                optionally returns the stage-1 smith_waterman paths
 *                                (caretta/dynamic_time_warping.py:225-278).
 *   crt_sw_align_batch        <- dtw.smith_waterman / smith_waterman_score on caller-supplied score matrices
 *                                (dynamic_time_warping.py:204-278)
 *   crt_dtw_align_batch       <- dtw.dtw_align / dtw_align_score                (dynamic_time_warping.py:7-201)
 *   crt_rmsd_cov_tm           <- make_rmsd_coverage_tm_matrix(superpose_first=False) (multiple_alignment.py:1000-1055)
 *
 * Conventions
 *   - plain pointers and sizes only; the caller (numpy) owns every host buffer, the library never keeps a host
 *     pointer past return and never frees caller memory.  Device buffers belong to the context.
 *   - every function returns 0 on success or a negative CRT_E_* code; crt_last_error() gives the message
 *     (thread-local).  Nothing throws or exits across the ABI.  There is no CPU fallback: without a CUDA
 *     device crt_create fails with CRT_E_CUDA.
 *   - data-dependent degeneracies are not errors; they set bits in out_status[pair]:
 *       CRT_ST_FEW_COMMON  <= 3 matched residues, superposition skipped exactly like multiple_alignment.py:337-342
 *       CRT_ST_NO_POSITIVE stage-1 score matrix had no cell > 0 (the reference raises from
 *                          dynamic_time_warping.py:250); the pair is scored without superposition
 *       CRT_ST_TIE         (CRT_FP32) the fp32 traceback met a decision the reference's float64 H matrix may take differently
 *                          (dynamic_time_warping.py:241-247, :260-277: increments below ulp(H) tie exactly there)
 *       CRT_ST_FP64        (CRT_FP32) such a pair was recomputed by the float64 kernels: its results are the CRT_FP64 ones
 *   - chain layout: coords[(offsets[p] + r) * 3 + axis], tensors[(offsets[p] + r) * d + k], float64, offsets[N+1].
 *   - precision: CRT_FP64 reproduces the reference's arithmetic (sequential non-fused RBF sum, equality
 *     traceback, row-major first maximum): stage-1 paths are identical to the reference's.  CRT_FP32 is the
 *     production mode (score / RMSD / TM within 1e-4 relative, >= 99.9 % identical aligned columns).
 *   - calls on one context are serialised by the caller; one context per (process, device).
 */
#ifndef CARETTA_B200_H
#define CARETTA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct crt_ctx crt_ctx;

enum { CRT_FP64 = 0, CRT_FP32 = 1 };

enum {
    CRT_OK = 0,
    CRT_E_ARG = -1,        /* bad argument (null pointer, negative size, unsupported d / gap / length) */
    CRT_E_CUDA = -2,       /* CUDA runtime failure (including: no device) */
    CRT_E_STATE = -3,      /* call order (e.g. pairwise before set_chains) */
    CRT_E_NOMEM = -4       /* device or host allocation failed */
};

enum { CRT_ST_FEW_COMMON = 1, CRT_ST_NO_POSITIVE = 2, CRT_ST_NONFINITE = 4, CRT_ST_TIE = 8, CRT_ST_FP64 = 16 };

/* Parameters of the pair recipe; defaults are the reference's (multiple_alignment.py:490-492, :335). */
typedef struct crt_params {
    double gamma_tensor;   /* 7.0  */
    double gamma_coords;   /* 0.03 */
    double sw_gap;         /* 0.0 -- the only value the reference uses on this path; others -> CRT_E_ARG */
    int32_t precision;     /* CRT_FP64 | CRT_FP32 */
    int32_t flags;         /* 0, or CRT_FLEXIBLE: Protein.score_function(flexible=True) (multiple_alignment.py:323-326): the score
                              matrix is the tensor Gaussian alone -> pair score = smith_waterman_score of it; rmsd / tm / ncommon 0 */
} crt_params;

enum { CRT_FLEXIBLE = 1 };
/* gamma_coords sentinels of the progressive-alignment calls (crt_progressive_node / _level, crt_msa_level) */
#define CRT_GAMMA_COORDS_FLEXIBLE (-1.0)        /* score_function(flexible=True) and mean_function(flexible=True) */
#define CRT_GAMMA_COORDS_FLEXIBLE_SCORE (-2.0)  /* score_function(flexible=True), mean_function(flexible=False) */

const char *crt_last_error(void);
int crt_version(void);

/* Context = one CUDA device + its streams and workspaces.  device < 0 -> current device. */
int crt_create(int device, crt_ctx **out);
int crt_destroy(crt_ctx *ctx);
int crt_device_info(crt_ctx *ctx, int32_t *sm_count, int32_t *clock_khz, int64_t *mem_bytes);

/* Upload (and preprocess on the device) the packed chain set.  Replaces any previous set. */
int crt_set_chains(crt_ctx *ctx, const double *coords, const double *tensors, const int64_t *offsets,
                   int32_t n_chains, int32_t d);

/* The same with coordinates only, for the consumers of an alignment (crt_superpose*, crt_rmsd_cov_tm*), which never read the
 * shape tensors; pair runs on such a chain set fail with CRT_E_STATE. */
int crt_set_coords(crt_ctx *ctx, const double *coords, const int64_t *offsets, int32_t n_chains);

/* All-vs-all, the shard of `rank` out of `world` (world = 1 -> everything).  Pairs are grouped into units
 * (one column chain x a run of row chains), units are dealt to ranks by cost; the enumeration is deterministic
 * so every rank can reconstruct every other rank's pair list with crt_shard_pairs.
 * Results stay in device memory (packed, in shard order) until fetched; the call returns after the work is
 * enqueued AND finished (synchronous), timing available from crt_last_elapsed_ms. */
int crt_pairwise_shard(crt_ctx *ctx, const crt_params *prm, int32_t rank, int32_t world);
int64_t crt_shard_size(crt_ctx *ctx, int32_t rank, int32_t world);      /* pairs in that shard, <0 on error */
int crt_shard_pairs(crt_ctx *ctx, int32_t rank, int32_t world, int32_t *pair_i, int32_t *pair_j);
/* The same enumeration without a context or a device (host-side planning: sizes of the all-gather, scatter maps). */
int64_t crt_plan_shard_size(const int64_t *offsets, int32_t n_chains, int32_t rank, int32_t world);
int crt_plan_shard_pairs(const int64_t *offsets, int32_t n_chains, int32_t rank, int32_t world, int32_t *pair_i,
                         int32_t *pair_j);
/* Copy the packed results of the last crt_pairwise_shard / crt_pairwise_list to host (any pointer may be NULL).
 * score/rmsd/tm are float64 at the boundary in both precisions (the reference returns float64). */
int crt_fetch(crt_ctx *ctx, double *score, double *rmsd, double *tm, int32_t *ncommon, int32_t *status);
/* Same, device to device: dst pointers are DEVICE addresses of float32 buffers (e.g. torch tensors used as NCCL
 * all-gather inputs), n = capacity in elements. */
int crt_fetch_device(crt_ctx *ctx, void *d_score, void *d_rmsd, void *d_tm, int64_t n);
double crt_last_elapsed_ms(crt_ctx *ctx);        /* device time of the last run (CUDA events on the run's stream) */
/* Per-phase device time of the last run: out4 = {stage-1 fill, traceback+Kabsch, (unused), stage-2 rows + fill}.
 * Only measured when the run used ONE stream (environment CARETTA_B200_STREAMS=1, or a path-returning call);
 * otherwise the phases of different batches overlap and every entry is -1. */
int crt_last_phase_ms(crt_ctx *ctx, double *out4);
int64_t crt_last_launches(crt_ctx *ctx);         /* kernels launched by the last run */
/* CRT_FP32 runs: how many pairs the tie detection sent through the float64 kernels, and the device time that took */
int crt_last_rerun(crt_ctx *ctx, int64_t *pairs, double *ms);
/* pairs of the last run whose stage-1 fill ran on the tensor-core kernel (environment CARETTA_B200_TC=1, experimental; 0 otherwise) */
int64_t crt_last_tc_pairs(crt_ctx *ctx);
double crt_last_cell_updates(crt_ctx *ctx);      /* sum over pairs of 2 * L1 * L2 */
/* bytes of traceback words the stage-1 fills of the last run wrote (from the allocation: strips x chunks x 32 lanes x 16 B per
 * unit) -- the HBM traffic of the dominant kernel */
double crt_last_traceback_bytes(crt_ctx *ctx);

/* ---- multi-GPU (the reference's seam is one process calling one method, multiple_alignment.py:498-500) ----
 * Packed exchange format of a shard: score | rmsd | tm, each `pad` elements (pad >= the largest shard), float32 in the
 * production mode or float64 (is_f64 = 1, exact) in the parity mode.  crt_pack_results fills that vector on the device
 * (d_dst is a DEVICE address, e.g. the input of an all-gather); crt_scatter_gathered takes the all-gathered
 * [world][3][pad] block (DEVICE address on the context's device), scatters it into the dense symmetric matrices on the
 * device and copies them to the caller's float64 [N,N] host arrays (out_rmsd / out_tm may be NULL). */
int crt_pack_results(crt_ctx *ctx, void *d_dst, int64_t pad, int32_t is_f64);
int crt_scatter_gathered(crt_ctx *ctx, const void *d_gathered, int32_t world, int64_t pad, int32_t is_f64,
                         double *out_score, double *out_rmsd, double *out_tm);
/* One process, several devices: a context per device and a NCCL communicator over them (ncclCommInitAll; NCCL is resolved with
 * dlopen, environment CARETTA_B200_NCCL overrides the library name).  n_dev <= 0: every visible device.
 * crt_multi_set_chains uploads the chains to every device (one host thread per device); crt_multi_pairwise_all =
 * make_pairwise_matrix over all of them: cost-sharded units, ONE grouped ncclAllGather of the packed vectors, dense scatter on
 * the first device, one copy to the host arrays.  Bitwise the one-GPU matrices. */
typedef struct crt_multi crt_multi;
int crt_multi_create(int32_t n_dev, const int32_t *dev_ids, crt_multi **out);
int crt_multi_destroy(crt_multi *m);
int32_t crt_multi_devices(crt_multi *m);
crt_ctx *crt_multi_ctx(crt_multi *m, int32_t index);
int crt_multi_set_chains(crt_multi *m, const double *coords, const double *tensors, const int64_t *offsets,
                         int32_t n_chains, int32_t d);
int crt_multi_pairwise_all(crt_multi *m, const crt_params *prm, double *out_score, double *out_rmsd, double *out_tm);
/* out3 = {max over devices of the shard's device ms, device ms of the all-gather, host wall ms of the call} */
int crt_multi_last_timing(crt_multi *m, double *out3, int64_t *rerun_pairs);

/* make_pairwise_matrix: dense symmetric float64 [N,N], diagonal 0 (rmsd/tm by-products: diagonal 0 / 1).
 * Convenience wrapper = crt_pairwise_shard(world=1) + crt_fetch + scatter. out_rmsd/out_tm may be NULL; when BOTH are NULL (what the
 * reference's method returns: the scores alone) the per-pair RMSD / TM by-products are not computed at all (crt_fetch reports 0). */
int crt_pairwise_all(crt_ctx *ctx, const crt_params *prm, double *out_score, double *out_rmsd, double *out_tm);
/* Page-locked host buffers for inputs/outputs of the calls above (optional; pageable memory works, slower). */
int crt_host_alloc(size_t bytes, void **out);
int crt_host_free(void *p);

/* Explicit pair list.  Optional stage-1 paths: aln_off[n_pairs+1] (int64, filled by the library), aln1/aln2 int32
 * with -1 = gap, ascending residue order like dynamic_time_warping.py:278; capacity aln_cap entries (sum over
 * pairs of L_i + L_j is always enough).  Pass NULL for all three to skip. */
int crt_pairwise_list(crt_ctx *ctx, const crt_params *prm, const int32_t *pair_i, const int32_t *pair_j,
                      int64_t n_pairs, double *score, double *rmsd, double *tm, int32_t *ncommon, int32_t *status,
                      int32_t *aln1, int32_t *aln2, int64_t *aln_off, int64_t aln_cap);

/* DP in isolation on caller-supplied float64 score matrices, problem p is S[shape_off[p] ...] with
 * n[p] x m[p] row-major.  aln1/aln2/aln_off as above (may be NULL for score only). */
int crt_sw_align_batch(crt_ctx *ctx, const double *S, const int64_t *shape_off, const int32_t *n, const int32_t *m,
                       int32_t n_problems, double gap, int32_t *aln1, int32_t *aln2, int64_t *aln_off,
                       int64_t aln_cap, double *score, int32_t *status);
int crt_dtw_align_batch(crt_ctx *ctx, const double *S, const int64_t *shape_off, const int32_t *n, const int32_t *m,
                        int32_t n_problems, double gap_open, double gap_extend, int32_t *aln1, int32_t *aln2,
                        int64_t *aln_off, int64_t aln_cap, double *score);

/* make_rmsd_coverage_tm_matrix(superpose_first=False) on the chains of the context: aln int64 [N, A], -1 = gap.
 * Outputs float64 [N,N] with the reference's diagonals (0 / 1 / 1).  *n_bad = pairs with < 3 common positions
 * (the reference asserts there); their entries keep the diagonal defaults. */
int crt_rmsd_cov_tm(crt_ctx *ctx, const int64_t *aln, int64_t A, double *rmsd, double *cov, double *tm,
                    int32_t *n_bad);
/* make_rmsd_coverage_tm_matrix(superpose_first=True) after crt_superpose: the same matrices on chains that are already in one
 * frame (no per-pair Kabsch, multiple_alignment.py:1025-1026, :1037). */
int crt_rmsd_cov_tm_superposed(crt_ctx *ctx, const int64_t *aln, int64_t A, double *rmsd, double *cov, double *tm,
                               int32_t *n_bad);

/* One node of the progressive alignment: replaces make_intermediate_node of MultipleAlignment.progressive_align
 * (multiple_alignment.py:195-234): score_matrix = Protein.score_function(n1, n2) (:321-349) + Gaussian of the scaled
 * consensus weights (:206-210), dtw_align with affine gaps (:211-214), Protein.mean_function (:351-383) and
 * get_mean_weights (:73-82).  All float64.  tensors: [n,d] / [m,d], coords: [n,3] / [m,3], weights: [n] / [m] (host).
 * Outputs: aln1/aln2 int32 with -1 = gap (capacity n + m), *aln_len, the intermediate node tensors_mean [len,d],
 * coords_mean [len,3], weights_mean [len]; *score = the DTW score; *status = status bits of the stage-1 pair run.
 * gamma_weight < 0 leaves the weight term out (multiple_align with two structures, :263-275).
 * gamma_coords < 0 selects flexible=True for this and the level / pool calls below (score_function :323-326): the score matrix is
 * the tensor Gaussian plus the weight term, no stage-1 alignment.  CRT_GAMMA_COORDS_FLEXIBLE: mean_function(flexible=True) too
 * (:359-360) -- no superposition, coords_mean is meaningless (the reference's flexible node has no coordinates);
 * CRT_GAMMA_COORDS_FLEXIBLE_SCORE: the node still superposes its children on the DTW alignment and averages coordinates. */
int crt_progressive_node(crt_ctx *ctx, const double *tensors1, const double *coords1, const double *weights1, int32_t n,
                         const double *tensors2, const double *coords2, const double *weights2, int32_t m, int32_t d,
                         double mult1, double mult2, double gamma_tensor, double gamma_coords, double gamma_weight,
                         double gap_open, double gap_extend, int32_t *aln1, int32_t *aln2, int32_t *aln_len,
                         double *tensors_mean, double *coords_mean, double *weights_mean, double *score, int32_t *status);

/* Protein.score_function (multiple_alignment.py:321-349) as the reference returns it: the full float64 [n,m] score matrix of
 * one pair (row-major, host).  flexible = 0: tensor Gaussian -> smith_waterman (gap 0) -> common positions -> Kabsch (skipped
 * when <= 3, *status = CRT_ST_FEW_COMMON) -> Gaussian of the superposed coordinates; flexible != 0 (:323-326): the tensor
 * Gaussian itself, coords1 / coords2 may be NULL.  Same kernels as the node path (crt_progressive_node without weight term). */
int crt_score_matrix(crt_ctx *ctx, const double *tensors1, const double *coords1, int32_t n, const double *tensors2,
                     const double *coords2, int32_t m, int32_t d, double gamma_tensor, double gamma_coords, int32_t flexible,
                     double *score_matrix, int32_t *status);

/* Protein.mean_function (multiple_alignment.py:351-383) for a given alignment (aln1 / aln2 int64 [len], -1 = gap, a column
 * with two gaps is CRT_E_ARG): tensors_mean [len,d] and -- flexible = 0 -- coords_mean [len,3] after superposing both chains on
 * the common positions of the alignment (no superposition when <= 3, *status = CRT_ST_FEW_COMMON).  flexible != 0 (:359-360):
 * tensors only, coords1 / coords2 / coords_mean may be NULL. */
int crt_mean_function(crt_ctx *ctx, const double *tensors1, const double *coords1, int32_t n, const double *tensors2,
                      const double *coords2, int32_t m, int32_t d, const int64_t *aln1, const int64_t *aln2, int64_t len,
                      int32_t flexible, double *tensors_mean, double *coords_mean, int32_t *status);

/* get_mean_weights (multiple_alignment.py:73-82): weights_mean[i] = (aln1[i] != -1 ? weights1[aln1[i]] : 0) +
 * (aln2[i] != -1 ? weights2[aln2[i]] : 0), float64 [len]. */
int crt_mean_weights(crt_ctx *ctx, const double *weights1, int32_t n, const double *weights2, int32_t m, const int64_t *aln1,
                     const int64_t *aln2, int64_t len, double *weights_mean);

/* All independent nodes of one level of the guide tree at once (nodes at the same depth are independent).
 * Node k aligns child chains 2k and 2k+1 of the packed level arrays: tensors [sum,d], coords [sum,3], weights [sum],
 * offsets [2 n_nodes + 1]; mult [n_nodes][2] = (multiplier_n1, multiplier_n2) of multiple_alignment.py:200-203.  Same computation
 * per node as crt_progressive_node.  Outputs are packed like the inputs: node k owns rows offsets[2k] .. offsets[2k] + aln_len[k]
 * (capacity n_k + m_k) of aln1 / aln2 (int32, -1 = gap), tensors_mean [sum,d], coords_mean [sum,3], weights_mean [sum];
 * score / status [n_nodes] (may be NULL). */
int crt_progressive_level(crt_ctx *ctx, int32_t n_nodes, int32_t d, const double *tensors, const double *coords, const double *weights,
                          const int64_t *offsets, const double *mult, double gamma_tensor, double gamma_coords, double gamma_weight,
                          double gap_open, double gap_extend, int32_t *aln1, int32_t *aln2, int32_t *aln_len, double *tensors_mean,
                          double *coords_mean, double *weights_mean, double *score, int32_t *status);

/* The whole progressive alignment with the sequences resident on the device (the loop of progressive_align,
 * multiple_alignment.py:236-249, level by level).  crt_msa_begin puts the chains of the context (crt_set_chains) into a sequence
 * pool as sequences 0..N-1 with consensus weight `consensus_weight` (:184-188).  crt_msa_level makes the n_nodes nodes
 * (child1[k], child2[k]) -- pool ids, any earlier sequences -- exactly like crt_progressive_level and appends them to the pool as
 * sequences *first_new_id + k; only the alignments come back: node k owns entries aln_off[k] .. aln_off[k] + aln_len[k] of
 * aln1 / aln2 (capacity aln_cap >= sum of the children's lengths).  crt_msa_lengths: number and lengths of the pool's sequences;
 * crt_msa_fetch: the listed sequences, packed (tensors [rows,d], coords [rows,3], weights [rows]); crt_msa_end frees the pool. */
int crt_msa_begin(crt_ctx *ctx, double consensus_weight, int32_t *n_sequences);
int crt_msa_level(crt_ctx *ctx, int32_t n_nodes, const int32_t *child1, const int32_t *child2, const double *mult, double gamma_tensor,
                  double gamma_coords, double gamma_weight, double gap_open, double gap_extend, int32_t *aln1, int32_t *aln2,
                  int64_t aln_cap, int64_t *aln_off, int32_t *aln_len, double *score, int32_t *status, int32_t *first_new_id);
int crt_msa_lengths(crt_ctx *ctx, int32_t *n_sequences, int32_t *lengths, int32_t cap);
int crt_msa_fetch(crt_ctx *ctx, const int32_t *ids, int32_t count, double *tensors, double *coords, double *weights);
/* crt_msa_compose: the index arrays (int64, -1 = gap) of the n_under sequences under pool sequence `root`, in the frame of `root`
 * (length A = its length): rows of out [n_under][A] in the reference's dictionary order (multiple_alignment.py:219-232: the first
 * child's sequences first), their pool ids in leaf_ids.  root = the last node gives progressive_align's return value (:253). */
int crt_msa_compose(crt_ctx *ctx, int32_t root, int32_t *leaf_ids, int32_t leaf_cap, int32_t *n_under, int64_t *out, int64_t out_cap);
int crt_msa_end(crt_ctx *ctx);

/* Neighbor joining on the device: replaces caretta/neighbor_joining.py:17-99 (`neighbor_joining(distance_matrix)`, called
 * at multiple_alignment.py:277 on max(S) - S of the pairwise matrix).  distance_matrix: float64 [N,N] row-major (host),
 * N >= 3.  tree: uint64 [2N-3][2] rows (node_1, node_2), node ids >= N are intermediate nodes in creation order;
 * branch_lengths: float64 [2N-3].  Bit-identical to the reference: sequential float64 row sums, Q in the reference's
 * operation order, first strict minimum in row-major order, new node at index 0.  crt_last_elapsed_ms = device time. */
int crt_neighbor_joining(crt_ctx *ctx, const double *distance_matrix, int32_t N, uint64_t *tree, double *branch_lengths,
                         int64_t *n_rows);

/* ---- consumers of the multiple alignment.  aln: int64 [N, A] row-major, -1 = gap, rows in the
 * order of the chains of the context (crt_set_chains) where coordinates are involved. ---- */

/* make_coverage_gap_distance_matrix (multiple_alignment.py:45-56), the input of get_reference_structures (:740-784):
 * distance[i][j] = (# columns where i has a residue and j a gap) / (# residues of i), aligning[i][j] = (# residues of i) minus
 * that count.  float64 [N,N] / int32 [N,N].  A protein without any residue in the alignment -> CRT_E_ARG (the reference divides
 * by zero).  Bit-identical to the reference (integer counts, one IEEE division).  Needs no chains. */
int crt_coverage_gap_matrix(crt_ctx *ctx, const int64_t *aln, int32_t N, int64_t A, double *distance, int32_t *aligning);

enum { CRT_SUP_AUTO = 0, CRT_SUP_CORE = 1, CRT_SUP_REFERENCE = 2 };

/* superpose (multiple_alignment.py:854-867): superpose_core (:869-905) when at least half of the columns are gap-free, else
 * superpose_reference (:908-927), on the coordinates of the context's chains.  mode: CRT_SUP_AUTO follows the reference's rule,
 * the other two force one of them.  reference < 0: the first protein with the most residues in the alignment (:855).
 * Pair p of the outputs is (reference, protein p): out_coords float64 [sum L, 3] (all chains, packed like the input),
 * out_rot [N,9] / out_tran [N,3] with x' = x R + t (apply_rotran, superposition_functions.py:63-80), out_ncommon [N] = columns
 * used.  Reference mode: a protein with <= 3 common columns keeps its coordinates (the reference asserts, :918); the caller sees it
 * in out_ncommon.  Any output except out_coords may be NULL. */
int crt_superpose(crt_ctx *ctx, const int64_t *aln, int64_t A, int32_t mode, int32_t reference, const int64_t *core_columns,
                  int64_t n_core_columns, double *out_coords, double *out_rot, double *out_tran, int32_t *out_ncommon,
                  int32_t *out_mode, int32_t *out_reference, int64_t *out_ncore);

/* The superposition loops of superpose_references (:930-950) and write_superposed_pdbs_reference(s) (:684-737, :787-850): pair q
 * superposes chain mem[q] onto chain ref[q] over their common alignment columns (helper.get_common_positions, helper.py:12-42) and
 * replaces mem[q]'s coordinates.  Batches [batch_off[b], batch_off[b+1]) run one after the other (a later batch sees the
 * coordinates an earlier one produced, like the reference's loop over its reference structures); the pairs of one batch run
 * concurrently, so inside a batch no chain may be both a member and a reference (a batch of one pair may superpose a chain onto
 * itself).  Pairs with <= 3 common columns are skipped (out_ncommon). */
int crt_superpose_pairs(crt_ctx *ctx, const int64_t *aln, int64_t A, const int32_t *ref, const int32_t *mem, int64_t n_pairs,
                        const int64_t *batch_off, int32_t n_batches, double *out_coords, double *out_rot, double *out_tran,
                        int32_t *out_ncommon);

/* helper.write_distance_matrix (helper.py:183-203): the text "n_rows\n" + for every row "name v v v ...\n" with each value
 * formatted like Python's f"{x:.4f}" (correctly rounded decimal expansion, ties to even; "nan", "inf", "-inf", "-0.0000").
 * matrix: float64 [n_rows, n_cols] (host), names: the row names as packed bytes, name_off [n_rows+1].  The text stays in device
 * memory; *out_len = its size in bytes, crt_text_fetch copies it out.  Byte-identical to the reference's file. */
int crt_format_matrix(crt_ctx *ctx, const double *matrix, int32_t n_rows, int32_t n_cols, const char *names, const int64_t *name_off,
                      int64_t *out_len);

/* MultipleAlignment.write_alignment / to_sequence_alignment (multiple_alignment.py:287-309): for every protein
 * ">name\n" + (sequence[aln[p][k]] or '-' for a gap, k < A) + "\n".  seqs / names: packed bytes with offsets [N+1].  An index
 * beyond its sequence -> CRT_E_ARG (the reference raises IndexError). */
int crt_format_fasta(crt_ctx *ctx, const int64_t *aln, int32_t N, int64_t A, const char *seqs, const int64_t *seq_off, const char *names,
                     const int64_t *name_off, int64_t *out_len);
int crt_text_fetch(crt_ctx *ctx, char *out, int64_t cap);        /* the text of the last crt_format_* call */

/* Fast-mode guide matrix (align_from_structure_files with full=False, multiple_alignment.py:503-511).
 * crt_count_matrix = make_count_matrix (:128-134): indices = the shapemer indices of all proteins packed, offsets [N+1];
 * out float64 [N, alphabet_size].  crt_braycurtis = braycurtis (:137-145): out[i][j] = sum_k |a_ik - b_jk| / sum_k |a_ik + b_jk|,
 * float64 [n1, n2], sums in k order (bit-identical to the numba loop; two all-zero rows give nan where the reference raises
 * ZeroDivisionError). */
int crt_count_matrix(crt_ctx *ctx, const int64_t *indices, const int64_t *offsets, int32_t N, int32_t alphabet_size, double *out);
int crt_braycurtis(crt_ctx *ctx, const double *counts_1, int32_t n1, const double *counts_2, int32_t n2, int32_t K, double *out);

/* FP32 FFMA micro-benchmark used as the measured roofline denominator: returns lane-FFMA/s. */
int crt_fp32_peak(crt_ctx *ctx, double *ffma_per_s, double *elapsed_ms);

#ifdef __cplusplus
}
#endif
#endif
