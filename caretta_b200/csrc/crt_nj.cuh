// crt_nj.cuh -- neighbor joining on the device (the immediate consumer of the pairwise matrix).
//
// Reference: caretta/neighbor_joining.py:17-157.  The reference recomputes np.sum(row) inside the O(n^2) scan of every
// iteration (O(N^4) overall) and rebuilds the matrix in Python every iteration; here one iteration is three launches over
// a dense n x n float64 matrix that ping-pongs between two buffers:
//   k_nj_argmin   Q[i][j] = ((n-2) * d[i][j] - sum_i) - sum_j in the reference's operation order, first strict minimum in
//                 row-major order (:118-129), block partials
//   k_nj_select   final reduction, branch lengths (:137-157), the two tree rows of the joined pair (:48-56)
//   k_nj_rebuild  new matrix with the joined node at index 0 and the remaining nodes in their previous order (:59-77),
//                 fused with the row sums of the NEW matrix
// Row sums are plain left-to-right float64 sums like numba's np.sum: a warp loads 32 consecutive elements coalesced and
// adds them in index order through shuffles, so the result is bit-identical to the sequential loop.
#pragma once
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdint.h>

namespace crt {

struct NjSel {
    double q;            // minimum of Q
    long long lin;       // i * n + j of the first row-major minimum
    int mi, mj;          // joined positions (current matrix)
    long long rows;      // tree rows written so far
    long long n_inter;   // intermediate nodes created so far
};

// acc + v(lane 0) + v(lane 1) + ... in lane order.  Full chunks are unrolled so that the 32 shuffles issue back to back and only
// the float64 adds form the dependent chain (the add order, hence the result, is the same as the rolled loop's).
__device__ __forceinline__ double nj_chain32(double acc, double v, int lim)
{
    if (lim == 32) {
#pragma unroll
        for (int l = 0; l < 32; ++l) acc = __dadd_rn(acc, __shfl_sync(0xffffffffu, v, l));
    } else {
        for (int l = 0; l < lim; ++l) acc = __dadd_rn(acc, __shfl_sync(0xffffffffu, v, l));
    }
    return acc;
}

// sequential (index-order) sum of one row by a warp; returns the sum in every lane
__device__ __forceinline__ double nj_row_sum(const double *row, int n, int lane)
{
    double acc = 0.0;
    double v = lane < n ? row[lane] : 0.0;
    for (int c0 = 0; c0 < n; c0 += 32) {
        const int c = c0 + 32 + lane;
        const double vn = c < n ? row[c] : 0.0;
        acc = nj_chain32(acc, v, min(32, n - c0));
        v = vn;
    }
    return acc;
}

__global__ void __launch_bounds__(256) k_nj_rowsums(const double *A, int n, double *S)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n) return;
    const double s = nj_row_sum(A + (size_t)warp * n, n, lane);
    if (lane == 0) S[warp] = s;
}

__device__ __forceinline__ bool nj_better(double q, long long lin, double bq, long long blin)
{
    return q < bq || (q == bq && lin < blin);
}

constexpr int NJ_ARGMIN_THREADS = 256, NJ_UNROLL = 4;

// Block b scans rows b, b + gridDim.x, ...; its threads stride over the columns (coalesced, no index division).  The minimum is
// taken over (q, linear index) pairs, so the traversal order does not matter: the winner is the first row-major minimum.
__device__ void nj_select_block(const double *A, const double *S, int n, int N, const double *pq, const long long *plin, int n_part,
                                const long long *true_idx, NjSel *sel, unsigned long long *tree, double *bl, double *sq, long long *sl);

// The last block to finish (ticket counter) also runs the selection, so that one iteration is two launches, not three.
// rows bid, bid + nb, ... of the scan, then the block reduction: the block's best (q, linear index) ends in sq[0] / sl[0]
__device__ __forceinline__ void nj_argmin_block(const double *A, const double *S, int n, int bid, int nb, double *sq, long long *sl)
{
    const double nm2 = (double)(n - 2);
    double bq = INFINITY;
    long long blin = 0;
    for (int i = bid; i < n; i += nb) {
        const double *row = A + (size_t)i * n;
        const double si = S[i];
        const long long base = (long long)i * n;
        int j = threadIdx.x;
        for (; j + (NJ_UNROLL - 1) * NJ_ARGMIN_THREADS < n; j += NJ_UNROLL * NJ_ARGMIN_THREADS) {   // independent loads in flight
            double a[NJ_UNROLL], sj[NJ_UNROLL];
#pragma unroll
            for (int u = 0; u < NJ_UNROLL; ++u) { a[u] = __ldcs(row + j + u * NJ_ARGMIN_THREADS); sj[u] = S[j + u * NJ_ARGMIN_THREADS]; }
#pragma unroll
            for (int u = 0; u < NJ_UNROLL; ++u) {
                const int jj = j + u * NJ_ARGMIN_THREADS;
                const double q = __dsub_rn(__dsub_rn(__dmul_rn(nm2, a[u]), si), sj[u]);
                if (jj != i && nj_better(q, base + jj, bq, blin)) { bq = q; blin = base + jj; }
            }
        }
        for (; j < n; j += NJ_ARGMIN_THREADS) {
            if (i == j) continue;
            const double q = __dsub_rn(__dsub_rn(__dmul_rn(nm2, row[j]), si), S[j]);
            if (nj_better(q, base + j, bq, blin)) { bq = q; blin = base + j; }
        }
    }
    sq[threadIdx.x] = bq; sl[threadIdx.x] = blin;
    __syncthreads();
    for (int w = NJ_ARGMIN_THREADS / 2; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w && nj_better(sq[threadIdx.x + w], sl[threadIdx.x + w], sq[threadIdx.x], sl[threadIdx.x])) {
            sq[threadIdx.x] = sq[threadIdx.x + w]; sl[threadIdx.x] = sl[threadIdx.x + w];
        }
        __syncthreads();
    }
}

// The last block to finish (ticket counter) also runs the selection, so that one iteration is two launches, not three.
__global__ void __launch_bounds__(NJ_ARGMIN_THREADS) k_nj_argmin(const double *A, const double *S, int n, double *pq, long long *plin,
                                                                 int N, const long long *true_idx, NjSel *sel, unsigned long long *tree,
                                                                 double *bl, unsigned *ticket)
{
    __shared__ double sq[NJ_ARGMIN_THREADS];
    __shared__ long long sl[NJ_ARGMIN_THREADS];
    nj_argmin_block(A, S, n, blockIdx.x, gridDim.x, sq, sl);
    __shared__ unsigned last;
    if (threadIdx.x == 0) {
        pq[blockIdx.x] = sq[0]; plin[blockIdx.x] = sl[0];
        __threadfence();
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last) {
        __threadfence();
        nj_select_block(A, S, n, N, pq, plin, (int)gridDim.x, true_idx, sel, tree, bl, sq, sl);
        if (threadIdx.x == 0) *ticket = 0;
    }
}

// one block: reduce the partials, write the two tree rows and branch lengths of the joined pair
__device__ void nj_select_block(const double *A, const double *S, int n, int N, const double *pq, const long long *plin, int n_part,
                                const long long *true_idx, NjSel *sel, unsigned long long *tree, double *bl, double *sq, long long *sl)
{
    double bq = INFINITY;
    long long blin = 0;
    for (int k = threadIdx.x; k < n_part; k += blockDim.x) {
        const double q = __ldcg(pq + k);
        const long long l = __ldcg(plin + k);
        if (nj_better(q, l, bq, blin)) { bq = q; blin = l; }
    }
    __syncthreads();
    sq[threadIdx.x] = bq; sl[threadIdx.x] = blin;
    __syncthreads();
    for (int w = NJ_ARGMIN_THREADS / 2; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w && nj_better(sq[threadIdx.x + w], sl[threadIdx.x + w], sq[threadIdx.x], sl[threadIdx.x])) {
            sq[threadIdx.x] = sq[threadIdx.x + w]; sl[threadIdx.x] = sl[threadIdx.x + w];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const long long lin = sl[0];
        const int mi = (int)(lin / n), mj = (int)(lin - (long long)mi * n);
        const double dij = A[(size_t)mi * n + mj];
        // _find_branch_length, neighbor_joining.py:137-157
        const double di = __dadd_rn(__dmul_rn(0.5, dij), __dmul_rn(0.5 / (double)(n - 2), __dsub_rn(S[mi], S[mj])));
        const double dj = __dsub_rn(dij, di);
        const long long node = sel->n_inter + N;
        long long r = sel->rows;
        tree[2 * r] = (unsigned long long)true_idx[mi]; tree[2 * r + 1] = (unsigned long long)node; bl[r] = di; ++r;
        tree[2 * r] = (unsigned long long)true_idx[mj]; tree[2 * r + 1] = (unsigned long long)node; bl[r] = dj; ++r;
        sel->q = sq[0]; sel->lin = lin; sel->mi = mi; sel->mj = mj; sel->rows = r; sel->n_inter += 1;
    }
}

// new (n-1) x (n-1) matrix B from A and the row sums of B.
//
// The row sums must be plain left-to-right float64 sums (numba's np.sum), i.e. one dependent add chain per row.  A CTA owns
// NJ_ROWS consecutive rows of B and walks them in super-tiles of NJ_TCOLS columns: all 8 warps gather the tile (coalesced 256-byte
// row segments of A -> B and into shared memory), then lane l of warp 0 adds the NJ_TCOLS values of row l in column order while
// every warp already has the loads of the next tile in flight (two shared-memory buffers).  One add per element instead of a
// 32-lane redundant chain with two shuffles per element: the kernel is bound by HBM (read n^2, write n^2), not by issue.
constexpr int NJ_ROWS = 16, NJ_TCOLS = 256, NJ_TSTRIDE = NJ_TCOLS + 1, NJ_REBUILD_THREADS = 256;
constexpr size_t NJ_REBUILD_SMEM = sizeof(double) * 2 * NJ_ROWS * NJ_TSTRIDE;

__device__ __forceinline__ void nj_rebuild_block(const double *A, int n, int mi, int mj, long long new_node, const long long *ti_old,
                                                 long long *ti_new, double *B, double *S_new, int row_block, double *nj_tile)
{
    const int nn = n - 1;
    const int r0 = row_block * NJ_ROWS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lo = min(mi, mj), hi = max(mi, mj);
    auto old_of = [&](int a) { int o = a; if (o >= lo) ++o; if (o >= hi) ++o; return o; };      // a-th remaining node -> old index
    const double dij = A[(size_t)mi * n + mj];
    const double *Ami = A + (size_t)mi * n, *Amj = A + (size_t)mj * n;
    const int rows = min(NJ_ROWS, nn - r0);
    // one super-tile: warp w takes the columns c0 + 32 w + lane of every row of the CTA.  fetch() only issues the loads (NJ_ROWS
    // independent 8-byte loads per thread stay in flight while warp 0 runs the add chain of the previous tile), put() stores.
    double v[NJ_ROWS];
    auto fetch = [&](int c0) {
        const int c = c0 + warp * 32 + lane;
        const int oc = c > 0 ? old_of(c - 1) : 0;
#pragma unroll
        for (int rr = 0; rr < NJ_ROWS; ++rr) {
            const int r = r0 + rr;
            v[rr] = 0.0;
            if (rr < rows && c < nn) {
                if (r == 0) {
                    if (c > 0) v[rr] = __dmul_rn(0.5, __dsub_rn(__dadd_rn(Ami[oc], Amj[oc]), dij));
                } else {
                    const int orow = old_of(r - 1);
                    v[rr] = c == 0 ? __dmul_rn(0.5, __dsub_rn(__dadd_rn(Ami[orow], Amj[orow]), dij)) : A[(size_t)orow * n + oc];
                }
            }
        }
    };
    auto put = [&](int c0, double *buf) {
        const int c = c0 + warp * 32 + lane;
#pragma unroll
        for (int rr = 0; rr < NJ_ROWS; ++rr) {
            if (rr < rows && c < nn) B[(size_t)(r0 + rr) * nn + c] = v[rr];
            buf[rr * NJ_TSTRIDE + warp * 32 + lane] = v[rr];
        }
    };
    const int n_tiles = (nn + NJ_TCOLS - 1) / NJ_TCOLS;
    double acc = 0.0;
    fetch(0);
    put(0, nj_tile);
    __syncthreads();
    for (int t = 0; t < n_tiles; ++t) {
        const double *cur = nj_tile + (size_t)(t & 1) * NJ_ROWS * NJ_TSTRIDE;
        const bool more = t + 1 < n_tiles;
        if (more) fetch((t + 1) * NJ_TCOLS);
        if (warp == 0 && lane < rows) {
            const int lim = min(NJ_TCOLS, nn - t * NJ_TCOLS);
            const double *rowv = cur + lane * NJ_TSTRIDE;
            if (lim == NJ_TCOLS) {
#pragma unroll 16
                for (int k = 0; k < NJ_TCOLS; ++k) acc = __dadd_rn(acc, rowv[k]);
            } else {
                for (int k = 0; k < lim; ++k) acc = __dadd_rn(acc, rowv[k]);
            }
        }
        if (more) put((t + 1) * NJ_TCOLS, nj_tile + (size_t)((t + 1) & 1) * NJ_ROWS * NJ_TSTRIDE);
        __syncthreads();
    }
    if (warp == 0 && lane < rows) {
        const int r = r0 + lane;
        S_new[r] = acc;
        ti_new[r] = r == 0 ? new_node : ti_old[old_of(r - 1)];
    }
}

__global__ void __launch_bounds__(NJ_REBUILD_THREADS) k_nj_rebuild(const double *A, int n, const NjSel *sel, int N, const long long *ti_old,
                                                                   long long *ti_new, double *B, double *S_new)
{
    extern __shared__ double nj_tile[];                 // [2][NJ_ROWS][NJ_TSTRIDE]
    nj_rebuild_block(A, n, sel->mi, sel->mj, sel->n_inter - 1 + N, ti_old, ti_new, B, S_new, blockIdx.x, nj_tile);
}

// ------------------------------------------------------------------------------------------------------------
// All iterations from n down to 4 in ONE cooperative launch (one CTA per SM, two grid-wide barriers per iteration instead of two
// dependent launches): below a few thousand nodes an iteration moves a few megabytes that sit in L2, and the ~14 us of launch
// turn-around were most of its 22 us.  Same device functions as the two-launch path, so the same bits: every CTA reduces the
// block partials itself (first row-major minimum), CTA 0 writes the two tree rows; the joined pair and the counters live in
// registers.  The host knows which buffer holds the final matrix from the number of iterations.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NJ_REBUILD_THREADS) k_nj_persistent(double *A, double *B, double *S0, double *S1, long long *t0, long long *t1,
                                                                      int n, int N, double *pq, long long *plin, NjSel *sel,
                                                                      unsigned long long *tree, double *bl)
{
    extern __shared__ double nj_tile[];                 // [2][NJ_ROWS][NJ_TSTRIDE]
    __shared__ double sq[NJ_ARGMIN_THREADS];
    __shared__ long long sl[NJ_ARGMIN_THREADS];
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    const int bid = blockIdx.x, nb = gridDim.x;
    long long rows = sel->rows, n_inter = sel->n_inter;
    while (n > 3) {
        nj_argmin_block(A, S0, n, bid, nb, sq, sl);
        if (threadIdx.x == 0) { pq[bid] = sq[0]; plin[bid] = sl[0]; }
        grid.sync();
        // every CTA: the first row-major minimum over the block partials
        double bq = INFINITY;
        long long blin = 0;
        for (int k = threadIdx.x; k < nb; k += blockDim.x) {
            const double q = __ldcg(pq + k);
            const long long l = __ldcg(plin + k);
            if (nj_better(q, l, bq, blin)) { bq = q; blin = l; }
        }
        __syncthreads();
        sq[threadIdx.x] = bq; sl[threadIdx.x] = blin;
        __syncthreads();
        for (int w = NJ_ARGMIN_THREADS / 2; w > 0; w >>= 1) {
            if ((int)threadIdx.x < w && nj_better(sq[threadIdx.x + w], sl[threadIdx.x + w], sq[threadIdx.x], sl[threadIdx.x])) {
                sq[threadIdx.x] = sq[threadIdx.x + w]; sl[threadIdx.x] = sl[threadIdx.x + w];
            }
            __syncthreads();
        }
        const long long lin = sl[0];
        const int mi = (int)(lin / n), mj = (int)(lin - (long long)mi * n);
        const long long node = n_inter + N;
        if (bid == 0 && threadIdx.x == 0) {
            const double dij = A[(size_t)mi * n + mj];
            // _find_branch_length, neighbor_joining.py:137-157
            const double di = __dadd_rn(__dmul_rn(0.5, dij), __dmul_rn(0.5 / (double)(n - 2), __dsub_rn(S0[mi], S0[mj])));
            const double dj = __dsub_rn(dij, di);
            tree[2 * rows] = (unsigned long long)t0[mi]; tree[2 * rows + 1] = (unsigned long long)node; bl[rows] = di;
            tree[2 * rows + 2] = (unsigned long long)t0[mj]; tree[2 * rows + 3] = (unsigned long long)node; bl[rows + 1] = dj;
        }
        __syncthreads();                                // sq / sl are reused by the next scan
        for (int rb = bid; rb * NJ_ROWS < n - 1; rb += nb) nj_rebuild_block(A, n, mi, mj, node, t0, t1, B, S1, rb, nj_tile);
        rows += 2; n_inter += 1;
        grid.sync();
        { double *x = A; A = B; B = x; x = S0; S0 = S1; S1 = x; long long *y = t0; t0 = t1; t1 = y; }
        --n;
    }
    if (bid == 0 && threadIdx.x == 0) { sel->rows = rows; sel->n_inter = n_inter; }
}

// the last three nodes, neighbor_joining.py:80-98 (n == 3)
__global__ void k_nj_last3(const double *A, const double *S, int N, const long long *true_idx, NjSel *sel, unsigned long long *tree, double *bl)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int n = 3;
    const double d12 = A[1 * n + 2];
    const double di = __dadd_rn(__dmul_rn(0.5, d12), __dmul_rn(0.5 / (double)(n - 2), __dsub_rn(S[1], S[2])));
    const double dj = __dsub_rn(d12, di);
    const long long node = sel->n_inter + N;
    long long r = sel->rows;
    tree[2 * r] = (unsigned long long)true_idx[1]; tree[2 * r + 1] = (unsigned long long)node; bl[r] = di; ++r;
    tree[2 * r] = (unsigned long long)true_idx[2]; tree[2 * r + 1] = (unsigned long long)node; bl[r] = dj; ++r;
    tree[2 * r] = (unsigned long long)true_idx[0]; tree[2 * r + 1] = (unsigned long long)node;
    bl[r] = __dmul_rn(0.5, __dsub_rn(__dadd_rn(A[1 * n + 0], A[2 * n + 0]), A[1 * n + 2])); ++r;
    sel->rows = r; sel->n_inter += 1;
}

// ------------------------------------------------------------------------------------------------------------
// IN-PLACE neighbor joining (round 2).  The two-buffer path above reads the matrix twice and WRITES it once per iteration
// (k_nj_rebuild moves every remaining row to put the new node at index 0).  Here nothing moves: the matrix is stored in
// REVERSED order (physical index p = logical index from the back), so "new node at logical index 0, the rest in their previous
// order" (neighbor_joining.py:59-77) is "append at the physical end"; the two joined nodes stay where they are as dead rows (never
// read again) and ZEROED columns (adding +0.0 leaves a left-to-right float64 sum bit for bit unchanged, so the row sums of numba's
// np.sum order come out of a plain descending sweep).  An iteration is
//   scan   Q over the alive rows, first row-major minimum of the LOGICAL order = largest (p, q) among equal values
//   join   every CTA: rows of the extent W + 1 in groups of NJ_ROWS, tiles of NJ_TCOLS columns from the back; the gather writes
//          the new column / row and the zeros of the two dead columns as it passes them, warp 0 runs the add chains
// = 2 n^2 reads and O(n) writes instead of 2 n^2 reads + n^2 writes, and the working set is ONE matrix (L2 holds it below ~3900
// nodes instead of ~2700).  The extent grows by one per join; the host compacts into the second buffer (same order) every n / 8
// joins.  Same arithmetic per element and per sum as the path above: bit-identical trees (tests/test_gpu_nj.py).
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool nj2_better(double q, long long key, double bq, long long bkey)
{
    return q < bq || (q == bq && key > bkey);
}

// block-wide winner (value, key): shuffles inside the warps, then one warp over the 8 warp winners; the result in every thread
__device__ __forceinline__ void nj2_block_best(double &bq, long long &bkey, double *sq, long long *sl)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double q = __shfl_xor_sync(0xffffffffu, bq, o);
        const long long k = __shfl_xor_sync(0xffffffffu, bkey, o);
        if (nj2_better(q, k, bq, bkey)) { bq = q; bkey = k; }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();                                    // sq / sl of an earlier use are no longer read
    if (lane == 0) { sq[warp] = bq; sl[warp] = bkey; }
    __syncthreads();
    constexpr int NW = NJ_ARGMIN_THREADS / 32;
    bq = lane < NW ? sq[lane] : INFINITY;
    bkey = lane < NW ? sl[lane] : -1;
#pragma unroll
    for (int o = NW / 2; o > 0; o >>= 1) {
        const double q = __shfl_xor_sync(0xffffffffu, bq, o);
        const long long k = __shfl_xor_sync(0xffffffffu, bkey, o);
        if (nj2_better(q, k, bq, bkey)) { bq = q; bkey = k; }
    }
    bq = __shfl_sync(0xffffffffu, bq, 0); bkey = __shfl_sync(0xffffffffu, bkey, 0);
}

// D[p][q] = in[N-1-p][N-1-q]
__global__ void __launch_bounds__(256) k_nj2_load(const double *in, int N, double *D, int ld, long long *tidx, int *alive, double *S_unused)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x, p = blockIdx.y;
    if (q >= N) return;
    D[(size_t)p * ld + q] = in[(size_t)(N - 1 - p) * N + (N - 1 - q)];
    if (p == 0) { tidx[q] = N - 1 - q; alive[q] = 1; }
}

// row sums in logical order (descending physical column) of a compact matrix: one warp per row, shuffle chain
__global__ void __launch_bounds__(256) k_nj2_rowsums(const double *D, int W, int ld, double *S)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= W) return;
    const double *row = D + (size_t)warp * ld;
    double acc = 0.0;
    for (int c0 = 0; c0 < W; c0 += 32) {
        const int k = c0 + lane;                           // position in the logical order
        const double v = k < W ? row[W - 1 - k] : 0.0;
        acc = nj_chain32(acc, v, min(32, W - c0));
    }
    if (lane == 0) S[warp] = acc;
}

constexpr int NJ2_RU = 4;                 // rows of the scan a thread keeps in flight
constexpr int NJ2_TC = 128, NJ2_TS = NJ2_TC + 2, NJ2_NBUF = 4;      // row stride even: 16-byte cp.async destinations
constexpr size_t NJ2_SMEM = sizeof(double) * NJ2_NBUF * NJ_ROWS * NJ2_TS;
__device__ __forceinline__ void cp_async16_nj(void *smem_dst, const void *gsrc)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit_nj() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_nj() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct Nj2Args {
    double *D; int ld;
    double *S; long long *tidx; int *alive;
    int n, W, iters, N;
    double *pq; long long *pkey;
    NjSel *sel; unsigned long long *tree; double *bl;
};

__global__ void __launch_bounds__(NJ_REBUILD_THREADS, 3) k_nj2_persistent(Nj2Args a)
{
    extern __shared__ double nj_tile[];                 // [NJ2_NBUF][NJ_ROWS][NJ2_TS]
    __shared__ double sq[NJ_ARGMIN_THREADS];
    __shared__ long long sl[NJ_ARGMIN_THREADS];
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    const int bid = blockIdx.x, nb = gridDim.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *D = a.D;
    const int ld = a.ld;
    int n = a.n, W = a.W;
    long long rows = a.sel->rows, n_inter = a.sel->n_inter;
    for (int it = 0; it < a.iters; ++it) {
        // ---- scan
        {
            const double nm2 = (double)(n - 2);
            double bq = INFINITY;
            long long bkey = -1;
            // The CTA's rows bid, bid + nb, ... are taken NJ2_RU at a time: a thread holds one pair of adjacent columns (16-byte
            // loads: ld and the column pairs are even) of NJ2_RU rows in flight and reads the column sums once for all of them --
            // rows one after the other would leave one load per thread in flight and pay a memory round trip per row.
            const double2 *S2 = reinterpret_cast<const double2 *>(a.S);
            const int W2 = W >> 1;
            for (int p0 = bid; p0 < W; p0 += NJ2_RU * nb) {
                const double2 *row2[NJ2_RU];
                double sp[NJ2_RU];
                int pr[NJ2_RU];
#pragma unroll
                for (int u = 0; u < NJ2_RU; ++u) {
                    const int p = p0 + u * nb;
                    const bool on = p < W && a.alive[p];
                    pr[u] = on ? p : -1;
                    row2[u] = reinterpret_cast<const double2 *>(D + (size_t)(on ? p : 0) * ld);
                    sp[u] = on ? a.S[p] : -INFINITY;          // a row that is not there: q = +inf everywhere
                }
                for (int c2 = threadIdx.x; c2 < W2; c2 += NJ_ARGMIN_THREADS) {
                    double2 v[NJ2_RU];
#pragma unroll
                    for (int u = 0; u < NJ2_RU; ++u) v[u] = row2[u][c2];
                    const double2 sc = S2[c2];
                    const int cc = 2 * c2;
#pragma unroll
                    for (int u = 0; u < NJ2_RU; ++u) {
                        const double q0 = __dsub_rn(__dsub_rn(__dmul_rn(nm2, v[u].x), sp[u]), sc.x);      // dead column: S = -inf, q = +inf
                        const double q1 = __dsub_rn(__dsub_rn(__dmul_rn(nm2, v[u].y), sp[u]), sc.y);
                        const long long base = (long long)pr[u] * ld;
                        if (cc != pr[u] && nj2_better(q0, base + cc, bq, bkey)) { bq = q0; bkey = base + cc; }
                        if (cc + 1 != pr[u] && nj2_better(q1, base + cc + 1, bq, bkey)) { bq = q1; bkey = base + cc + 1; }
                    }
                }
                if ((W & 1) && threadIdx.x == 0) {          // the odd last column
                    const int cc = W - 1;
#pragma unroll
                    for (int u = 0; u < NJ2_RU; ++u) {
                        if (pr[u] < 0 || pr[u] == cc) continue;
                        const double q = __dsub_rn(__dsub_rn(__dmul_rn(nm2, D[(size_t)pr[u] * ld + cc]), sp[u]), a.S[cc]);
                        if (nj2_better(q, (long long)pr[u] * ld + cc, bq, bkey)) { bq = q; bkey = (long long)pr[u] * ld + cc; }
                    }
                }
            }
            nj2_block_best(bq, bkey, sq, sl);
            if (threadIdx.x == 0) { a.pq[bid] = bq; a.pkey[bid] = bkey; }
        }
        grid.sync();
        // ---- every CTA: the winner over the block partials
        long long key;
        {
            double bq = INFINITY;
            long long bkey = -1;
            for (int k = threadIdx.x; k < nb; k += blockDim.x) {
                const double q = __ldcg(a.pq + k);
                const long long l = __ldcg(a.pkey + k);
                if (nj2_better(q, l, bq, bkey)) { bq = q; bkey = l; }
            }
            nj2_block_best(bq, bkey, sq, sl);
            key = bkey;
        }
        const int pi = (int)(key / ld), pj = (int)(key - (long long)pi * ld);       // row = the reference's min_i, column = min_j
        const long long node = n_inter + a.N;
        const double dij = D[(size_t)pi * ld + pj];
        if (bid == 0 && threadIdx.x == 0) {
            // _find_branch_length, neighbor_joining.py:137-157
            const double di = __dadd_rn(__dmul_rn(0.5, dij), __dmul_rn(0.5 / (double)(n - 2), __dsub_rn(a.S[pi], a.S[pj])));
            const double dj = __dsub_rn(dij, di);
            a.tree[2 * rows] = (unsigned long long)a.tidx[pi]; a.tree[2 * rows + 1] = (unsigned long long)node; a.bl[rows] = di;
            a.tree[2 * rows + 2] = (unsigned long long)a.tidx[pj]; a.tree[2 * rows + 3] = (unsigned long long)node; a.bl[rows + 1] = dj;
            // the two joined nodes die, the new one takes the end of the extent (nobody reads these entries before the next barrier:
            // the join below names pi, pj and W explicitly)
            a.S[pi] = -INFINITY; a.S[pj] = -INFINITY;
            a.alive[pi] = 0; a.alive[pj] = 0; a.alive[W] = 1;
            a.tidx[W] = node;
        }
        // ---- join + row sums of the new matrix.  Per CTA and group of NJ_ROWS rows: warps 1..4 stream the row segments into a
        //      ring of NJ2_NBUF tiles with cp.async (no register staging: NJ2_NBUF - 1 tiles stay in flight), warp 0 patches the
        //      three special columns of a tile (new column, the two joined ones) and runs the add chains -- the chain of n
        //      dependent float64 adds per row is what an iteration of this pass costs, the loads hide under it
        {
            const double *Dpi = D + (size_t)pi * ld, *Dpj = D + (size_t)pj * ld;
            const int We = W + 1;                           // extent with the new node at physical index W
            const int n_tiles = (We + NJ2_TC - 1) / NJ2_TC; // tile tau = columns [128 tau, 128 tau + 128), walked from the last to the first
            for (int rb = bid; rb * NJ_ROWS < We; rb += nb) {
                const int r0 = rb * NJ_ROWS;
                const int prow = r0 + lane;                 // warp 0: the row whose add chain this lane runs
                double jpa = 0.0, jpb = 0.0;                // (loaded before the flags below are waited for: one round trip, not two)
                if (warp == 0 && lane < NJ_ROWS) { jpa = Dpi[prow]; jpb = Dpj[prow]; }
                unsigned act = 0;                           // rows of the group that belong to the new matrix
#pragma unroll
                for (int rr = 0; rr < NJ_ROWS; ++rr) {
                    const int p = r0 + rr;
                    if (p == W || (p < W && p != pi && p != pj && a.alive[p])) act |= 1u << rr;
                }
                if (act == 0) continue;
                const int rr_new = W - r0;                  // the new node's row, if 0 <= rr_new < NJ_ROWS
                // gather thread g = 0..127 (warps 1..4): columns 2 (g & 63), +1 of the tile, rows 8 (g >> 6) .. + 7 of the group
                const int g = (int)threadIdx.x - 32;
                auto issue = [&](int t) {                   // t-th tile of the walk = tile tau = n_tiles - 1 - t
                    if (g < 0 || g >= 128 || t >= n_tiles) return;
                    double *buf = nj_tile + (size_t)(t % NJ2_NBUF) * NJ_ROWS * NJ2_TS;
                    const int k = 2 * (g & 63), c = (n_tiles - 1 - t) * NJ2_TC + k;
                    if (c >= We) return;                    // (c + 1 == We is read and not used: the row has ld >= We + 1 elements)
#pragma unroll
                    for (int q = 0; q < NJ_ROWS / 2; ++q) {
                        const int rr = (g >> 6) * (NJ_ROWS / 2) + q;
                        if (!((act >> rr) & 1u)) continue;
                        double *dst = buf + rr * NJ2_TS + k;
                        if (rr == rr_new) {
                            // the new node's row: 0 at itself and at dead columns (neighbor_joining.py:59-77)
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                const int cc = c + h;
                                double jc = 0.0;
                                if (cc < W && cc != pi && cc != pj && a.alive[cc]) jc = __dmul_rn(0.5, __dsub_rn(__dadd_rn(Dpi[cc], Dpj[cc]), dij));
                                dst[h] = jc;
                                if (cc < We) D[(size_t)W * ld + cc] = jc;
                            }
                        } else {
                            cp_async16_nj(dst, D + (size_t)(r0 + rr) * ld + c);
                        }
                    }
                };
                double acc = 0.0, jp = 0.0;                 // jp: the new node's distance to this lane's row
                const bool chain = warp == 0 && lane < NJ_ROWS && ((act >> lane) & 1u);
                if (chain && lane != rr_new) jp = __dmul_rn(0.5, __dsub_rn(__dadd_rn(jpa, jpb), dij));
                __syncthreads();                            // the ring is free (previous group)
                for (int t = 0; t < NJ2_NBUF - 1; ++t) { issue(t); cp_async_commit_nj(); }
                for (int t = 0; t < n_tiles; ++t) {
                    cp_async_wait_nj<NJ2_NBUF - 2>();        // this thread's copies of tile t have landed
                    __syncthreads();                        // everybody's have, and warp 0 is through with tile t - 1
                    issue(t + NJ2_NBUF - 1);                // into the buffer of tile t - 1
                    cp_async_commit_nj();
                    if (chain) {
                        double *rowv = nj_tile + (size_t)(t % NJ2_NBUF) * NJ_ROWS * NJ2_TS + lane * NJ2_TS;
                        const int c0 = (n_tiles - 1 - t) * NJ2_TC;      // column of position 0
                        if (lane != rr_new) {
                            // special columns of an ordinary row: the new column (the join value), the joined nodes' columns (zero)
                            const int kw = W - c0, ki = pi - c0, kj = pj - c0;
                            if (kw >= 0 && kw < NJ2_TC) { rowv[kw] = jp; D[(size_t)prow * ld + W] = jp; }
                            if (ki >= 0 && ki < NJ2_TC) { rowv[ki] = 0.0; D[(size_t)prow * ld + pi] = 0.0; }
                            if (kj >= 0 && kj < NJ2_TC) { rowv[kj] = 0.0; D[(size_t)prow * ld + pj] = 0.0; }
                        }
                        // logical order = descending physical column
                        const int hi = min(NJ2_TC, We - c0) - 1;
                        if (hi == NJ2_TC - 1) {
#pragma unroll 16
                            for (int k = NJ2_TC - 1; k >= 0; --k) acc = __dadd_rn(acc, rowv[k]);
                        } else {
                            for (int k = hi; k >= 0; --k) acc = __dadd_rn(acc, rowv[k]);
                        }
                    }
                }
                cp_async_wait_nj<0>();
                if (chain) a.S[prow] = acc;
            }
        }
        grid.sync();
        rows += 2; n_inter += 1; --n; ++W;
    }
    if (bid == 0 && threadIdx.x == 0) { a.sel->rows = rows; a.sel->n_inter = n_inter; }
}

// compaction: map[p'] = p-th alive physical index (one block), then a gather of rows and columns in the same order
__global__ void __launch_bounds__(1024) k_nj2_map(const int *alive, int W, int *map, int *count)
{
    __shared__ int wsum[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int c0 = 0; c0 < W; c0 += 1024) {
        const int p = c0 + threadIdx.x;
        const int f = p < W && alive[p] ? 1 : 0;
        const unsigned bal = __ballot_sync(0xffffffffu, f);
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        if (lane == 0) wsum[w] = __popc(bal);
        __syncthreads();
        int before = carry;
        for (int k = 0; k < w; ++k) before += wsum[k];
        if (f) map[before + __popc(bal & ((1u << lane) - 1u))] = p;
        __syncthreads();
        if (threadIdx.x == 0) { int s = 0; for (int k = 0; k < 32; ++k) s += wsum[k]; carry += s; }
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = carry;
}
__global__ void __launch_bounds__(256) k_nj2_compact(const double *Ds, int lds, const int *map, int n, double *Dd, int ldd, const double *Ss,
                                                     double *Sd, const long long *ts, long long *td, int *alive_d)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x, p = blockIdx.y;
    if (q >= n) return;
    Dd[(size_t)p * ldd + q] = Ds[(size_t)map[p] * lds + map[q]];
    if (p == 0) { Sd[q] = Ss[map[q]]; td[q] = ts[map[q]]; alive_d[q] = 1; }
}

// the last three nodes (neighbor_joining.py:80-98) on a compact reversed matrix: logical index a = physical 2 - a
__global__ void k_nj2_last3(const double *D, int ld, const double *S, int N, const long long *tidx, NjSel *sel, unsigned long long *tree, double *bl)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int n = 3;
    auto A = [&](int i, int j) { return D[(size_t)(2 - i) * ld + (2 - j)]; };
    const double d12 = A(1, 2);
    const double di = __dadd_rn(__dmul_rn(0.5, d12), __dmul_rn(0.5 / (double)(n - 2), __dsub_rn(S[2 - 1], S[2 - 2])));
    const double dj = __dsub_rn(d12, di);
    const long long node = sel->n_inter + N;
    long long r = sel->rows;
    tree[2 * r] = (unsigned long long)tidx[2 - 1]; tree[2 * r + 1] = (unsigned long long)node; bl[r] = di; ++r;
    tree[2 * r] = (unsigned long long)tidx[2 - 2]; tree[2 * r + 1] = (unsigned long long)node; bl[r] = dj; ++r;
    tree[2 * r] = (unsigned long long)tidx[2 - 0]; tree[2 * r + 1] = (unsigned long long)node;
    bl[r] = __dmul_rn(0.5, __dsub_rn(__dadd_rn(A(1, 0), A(2, 0)), A(1, 2))); ++r;
    sel->rows = r; sel->n_inter += 1;
}

}  // namespace crt
