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
#include <stdint.h>

namespace crt {

struct NjSel {
    double q;            // minimum of Q
    long long lin;       // i * n + j of the first row-major minimum
    int mi, mj;          // joined positions (current matrix)
    long long rows;      // tree rows written so far
    long long n_inter;   // intermediate nodes created so far
};

// sequential (index-order) sum of one row by a warp; returns the sum in every lane
__device__ __forceinline__ double nj_row_sum(const double *row, int n, int lane)
{
    double acc = 0.0;
    for (int c0 = 0; c0 < n; c0 += 32) {
        const int c = c0 + lane;
        const double v = c < n ? row[c] : 0.0;
        const int lim = min(32, n - c0);
        for (int l = 0; l < lim; ++l) acc = __dadd_rn(acc, __shfl_sync(0xffffffffu, v, l));
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

constexpr int NJ_ARGMIN_THREADS = 256;

__global__ void __launch_bounds__(NJ_ARGMIN_THREADS) k_nj_argmin(const double *A, const double *S, int n, double *pq, long long *plin)
{
    const long long total = (long long)n * n;
    const double nm2 = (double)(n - 2);
    double bq = INFINITY;
    long long blin = 0;
    for (long long lin = (long long)blockIdx.x * blockDim.x + threadIdx.x; lin < total; lin += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(lin / n), j = (int)(lin - (long long)i * n);
        if (i == j) continue;
        const double q = __dsub_rn(__dsub_rn(__dmul_rn(nm2, A[lin]), S[i]), S[j]);
        if (nj_better(q, lin, bq, blin)) { bq = q; blin = lin; }
    }
    __shared__ double sq[NJ_ARGMIN_THREADS];
    __shared__ long long sl[NJ_ARGMIN_THREADS];
    sq[threadIdx.x] = bq; sl[threadIdx.x] = blin;
    __syncthreads();
    for (int w = NJ_ARGMIN_THREADS / 2; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w && nj_better(sq[threadIdx.x + w], sl[threadIdx.x + w], sq[threadIdx.x], sl[threadIdx.x])) {
            sq[threadIdx.x] = sq[threadIdx.x + w]; sl[threadIdx.x] = sl[threadIdx.x + w];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { pq[blockIdx.x] = sq[0]; plin[blockIdx.x] = sl[0]; }
}

// one block: reduce the partials, write the two tree rows and branch lengths of the joined pair
__global__ void __launch_bounds__(NJ_ARGMIN_THREADS) k_nj_select(const double *A, const double *S, int n, int N, const double *pq,
                                                                 const long long *plin, int n_part, const long long *true_idx,
                                                                 NjSel *sel, unsigned long long *tree, double *bl)
{
    __shared__ double sq[NJ_ARGMIN_THREADS];
    __shared__ long long sl[NJ_ARGMIN_THREADS];
    double bq = INFINITY;
    long long blin = 0;
    for (int k = threadIdx.x; k < n_part; k += blockDim.x)
        if (nj_better(pq[k], plin[k], bq, blin)) { bq = pq[k]; blin = plin[k]; }
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

// new (n-1) x (n-1) matrix B from A and the row sums of B; one warp per row of B
__global__ void __launch_bounds__(256) k_nj_rebuild(const double *A, int n, const NjSel *sel, int N, const long long *ti_old, long long *ti_new,
                                                    double *B, double *S_new)
{
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nn = n - 1;
    if (r >= nn) return;
    const int mi = sel->mi, mj = sel->mj;
    const int lo = min(mi, mj), hi = max(mi, mj);
    auto old_of = [&](int a) { int o = a; if (o >= lo) ++o; if (o >= hi) ++o; return o; };      // a-th remaining node -> old index
    const double dij = A[(size_t)mi * n + mj];
    const double *Ami = A + (size_t)mi * n, *Amj = A + (size_t)mj * n;
    const int orow = r > 0 ? old_of(r - 1) : 0;
    const double *Arow = A + (size_t)orow * n;
    double *Brow = B + (size_t)r * nn;
    double acc = 0.0;
    for (int c0 = 0; c0 < nn; c0 += 32) {
        const int c = c0 + lane;
        double v = 0.0;
        if (c < nn) {
            if (r == 0) {
                if (c > 0) { const int oc = old_of(c - 1); v = __dmul_rn(0.5, __dsub_rn(__dadd_rn(Ami[oc], Amj[oc]), dij)); }
            } else if (c == 0) {
                v = __dmul_rn(0.5, __dsub_rn(__dadd_rn(Ami[orow], Amj[orow]), dij));
            } else {
                v = Arow[old_of(c - 1)];
            }
            Brow[c] = v;
        }
        const int lim = min(32, nn - c0);
        for (int l = 0; l < lim; ++l) acc = __dadd_rn(acc, __shfl_sync(0xffffffffu, v, l));
    }
    if (lane == 0) {
        S_new[r] = acc;
        ti_new[r] = r == 0 ? (sel->n_inter - 1 + N) : ti_old[orow];
    }
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

}  // namespace crt
