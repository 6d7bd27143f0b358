// crt_dp_batch.cuh -- the DPs of the path in isolation, on caller-supplied fp64 score matrices, plus the
// MSA-column RMSD / coverage / TM kernel.  Everything here is fp64 and reproduces the reference's decisions exactly:
//   k_dtw_fill / k_dtw_trace   dynamic_time_warping.py:7-86 (_make_dtw_matrix), :89-144 (_get_dtw_alignment), :147-201
//   k_sw_fill / k_sw_trace     dynamic_time_warping.py:204-278 (any gap >= 0: H is stored, the traceback re-evaluates
//                              the reference's equalities literally)
//   k_rmsd_cov_tm              multiple_alignment.py:1000-1055 (superpose_first=False), :59-70, score_functions.py:14-19
// Same systolic layout as the pair kernels: one warp per problem, lane = DPC consecutive columns, strips of 32*DPC
// columns with the strip boundary column kept in global memory.
#pragma once
#include "crt_kernels.cuh"
#include <cfloat>
#include <cstdlib>

namespace crt {

constexpr int DPC = 4;                // columns per lane
constexpr int DPSTRIP = 32 * DPC;
constexpr int DTW_PF = 8;             // wavefront steps of prefetch distance in k_dtw_fill (power of two)

struct DpProblem {
    long long s_off;      // offset of S (doubles)
    long long b_off;      // offset into the byte / H workspace (cells)
    long long bnd_off;    // offset into the boundary workspace (rows)
    long long aln_off;    // offset into the alignment output (capacity n + m + 1)
    int n, m;
};

__device__ __forceinline__ double shfl_up_d(double v) { return __shfl_up_sync(FULL, v, 1); }

// Row pitch of the affine DP's backtrack bytes: rows start on a 4-byte boundary so that a lane stores its DPC = 4 codes of a row as
// one 32-bit word (the fill is bound by the number of memory transactions per wavefront step: every lane is on a different row).
__host__ __device__ __forceinline__ int dtw_pitch(int m) { return (m + 3) & ~3; }
static_assert(DPC == 4, "k_dtw_fill packs the DPC backtrack codes of a lane into one 32-bit store");

// ------------------------------------------------------------------------------------------------------------
// Affine three-state DP.  States: 0 = lower (consumes i), 1 = match, 2 = upper (consumes j); ties -> lowest index
// (np.argmax).  One byte per cell: bit0 = B[.,.,0], bits1-2 = B[.,.,1], bit3 = B[.,.,2] - 1.
// ------------------------------------------------------------------------------------------------------------
// Scores and (strips > 0) the boundary column do not depend on the recurrence: they are fetched DTW_PF wavefront steps ahead
// with cp.async into a ring in shared memory (one 8-byte slot per lane and column), so that the loop body stays ONE step long.
// (A register ring needs the loop unrolled by the ring size; at four steps the body was 22 KB of code and the kernel ran at the
// instruction-fetch rate -- measured: half the unroll, 20 % faster; three warps of one CTA on different steps, 3.7x slower each.)
__device__ __forceinline__ void dtw_cp_async8(void *smem_dst, const void *gmem_src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void dtw_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void dtw_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(32) k_dtw_fill(const DpProblem *probs, int n_probs, const double *S_all, unsigned char *B_all,
                                                 double *bnd_all, double *final3, double open, double ext)
{
    __shared__ double ring[DTW_PF][DPC + 2][32];                      // [step slot][column 0..3, boundary 1, boundary 2][lane]
    if ((int)blockIdx.x >= n_probs) return;
    const DpProblem pr = probs[blockIdx.x];
    const int lane = threadIdx.x, n = pr.n, m = pr.m;
    const double *S = S_all + pr.s_off;
    unsigned char *B = B_all + pr.b_off;                              // [n][dtw_pitch(m)]
    const int mp = dtw_pitch(m);
    double *bnd1 = bnd_all + pr.bnd_off * 2, *bnd2 = bnd1 + n;        // M[i][cend][1], M[i][cend][2] per row i (1-based -> i-1)
    const double MINF = -DBL_MAX;
    const int n_strips = (m + DPSTRIP - 1) / DPSTRIP;
    for (int q = 0; q < DTW_PF * (DPC + 2); ++q) (&ring[0][0][0])[q * 32 + lane] = 0.0;     // columns past m are never fetched
    __syncwarp();
    for (int strip = 0; strip < n_strips; ++strip) {
        const int c0 = strip * DPSTRIP + lane * DPC;                 // 0-based first owned column
        double P0[DPC], P1[DPC];
#pragma unroll
        for (int c = 0; c < DPC; ++c) { P0[c] = MINF - open; P1[c] = 0.0; }      // row 0: (MIN - open, 0, 0)
        double out1 = 0.0, out2 = 0.0;       // M[i][cend][1], [2] handed to lane + 1
        double dsave = 0.0;                  // M[i-1][c0-1][1]
        const bool last_strip = strip == n_strips - 1;
        auto prefetch = [&](int t) {         // one (possibly empty) cp.async group per wavefront step
            const int i = t - lane + 1;
            if (i >= 1 && i <= n) {
                double (*slot)[32] = ring[t & (DTW_PF - 1)];
                const double *src = S + (long long)(i - 1) * m + c0;
#pragma unroll
                for (int c = 0; c < DPC; ++c)
                    if (c0 + c < m) dtw_cp_async8(&slot[c][lane], src + c);
                if (lane == 0 && strip > 0) { dtw_cp_async8(&slot[DPC][0], bnd1 + i - 1); dtw_cp_async8(&slot[DPC + 1][0], bnd2 + i - 1); }
            }
            dtw_cp_commit();
        };
        for (int u = 0; u < DTW_PF; ++u) prefetch(u);
        const int T = n + 31;
        for (int t = 0; t < T; ++t) {
            const int i = t - lane + 1;      // 1-based row
            const bool valid = i >= 1 && i <= n;
            dtw_cp_wait<DTW_PF - 1>();     // the group of step t has landed (DTW_PF - 1 younger groups may be in flight)
            double (*slot)[32] = ring[t & (DTW_PF - 1)];
            double sc[DPC];
#pragma unroll
            for (int c = 0; c < DPC; ++c) sc[c] = slot[c][lane];
            double b1 = 0.0, b2 = 0.0;        // boundary values: lane 0 fetched them, lane 0 alone reads them
            if (lane == 0) { b1 = slot[DPC][0]; b2 = slot[DPC + 1][0]; }
            prefetch(t + DTW_PF);            // refills this slot; every lane has read its entries above
            double L1 = shfl_up_d(out1), L2 = shfl_up_d(out2);
            if (lane == 0) {
                if (strip == 0) { L1 = 0.0; L2 = MINF - open; }       // column 0: (0, 0, MIN - open)
                else if (valid) { L1 = b1; L2 = b2; }
            }
            if (i == 1) dsave = 0.0;          // M[0][j][1] = 0
            const double in1 = L1;
            double D1 = dsave;
            if (valid) {
                unsigned codes = 0;
#pragma unroll
                for (int c = 0; c < DPC; ++c) {
                    const int j = c0 + c;     // 0-based column
                    const double s = sc[c];
                    const double l0 = P0[c] - ext, l1 = P1[c] - open;
                    const int ql = l1 > l0 ? 1 : 0;
                    const double lower = ql ? l1 : l0;
                    const double u0 = L1 - open, u1 = L2 - ext;
                    const int qu = u1 > u0 ? 1 : 0;
                    const double upper = qu ? u1 : u0;
                    const double dg = D1 + s;
                    double v = lower; int q = 0;
                    if (dg > v) { v = dg; q = 1; }
                    if (upper > v) { v = upper; q = 2; }
                    codes |= (unsigned)(ql | (q << 1) | (qu << 3)) << (8 * c);
                    if (i == n && j == m - 1) { final3[blockIdx.x * 3] = lower; final3[blockIdx.x * 3 + 1] = v; final3[blockIdx.x * 3 + 2] = upper; }
                    D1 = P1[c];
                    P0[c] = lower; P1[c] = v;
                    L1 = v; L2 = upper;
                }
                if (c0 < m) *reinterpret_cast<unsigned *>(B + (long long)(i - 1) * mp + c0) = codes;     // bytes past column m - 1: padding
                out1 = L1; out2 = L2;
                dsave = in1;
                if (!last_strip && lane == 31) { bnd1[i - 1] = out1; bnd2[i - 1] = out2; }
            }
        }
        dtw_cp_wait<0>();
        __syncwarp();
    }
}

// The strips of a problem on concurrent warps (the scheme of k_fill_s64_mw, crt_node_fill.cuh): warp w runs strip s0 + w,
// DTW_SKEW steps behind warp w - 1, all warps in lockstep with one __syncthreads per four steps; lane 31 of a warp hands
// M[i][cend][1], [2] to lane 0 of the next through a 64-row ring in shared memory (written at global step i + 30 + DTW_SKEW w, read
// at i - 1 + DTW_SKEW (w + 1): five steps later, behind a barrier).  n + 31 + 36 (strips - 1) steps instead of (n + 31) strips; same
// operands in the same order per cell, so the same codes and final values.  Used for tree levels of a few nodes, where the
// level waits for single warps; rounds of NW strips when a problem has more strips than the CTA has warps.
constexpr int DTW_SKEW = 36;
constexpr int DTW_MW_MAX = 12;
__global__ void __launch_bounds__(32 * DTW_MW_MAX) k_dtw_fill_mw(const DpProblem *probs, int n_probs, const double *S_all, unsigned char *B_all,
                                                                 double *bnd_all, double *final3, double open, double ext, int NW)
{
    extern __shared__ double dtw_dyn[];
    if ((int)blockIdx.x >= n_probs) return;
    const DpProblem pr = probs[blockIdx.x];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n = pr.n, m = pr.m;
    double (*ring)[DPC + 2][32] = reinterpret_cast<double (*)[DPC + 2][32]>(dtw_dyn + (size_t)warp * DTW_PF * (DPC + 2) * 32);
    double *xch = dtw_dyn + (size_t)NW * DTW_PF * (DPC + 2) * 32;     // [NW][64][2]
    const double *S = S_all + pr.s_off;
    unsigned char *B = B_all + pr.b_off;                              // [n][dtw_pitch(m)]
    const int mp = dtw_pitch(m);
    double *bnd1 = bnd_all + pr.bnd_off * 2, *bnd2 = bnd1 + n;
    const double MINF = -DBL_MAX;
    const int n_strips = (m + DPSTRIP - 1) / DPSTRIP;
    for (int q = 0; q < DTW_PF * (DPC + 2); ++q) (&ring[0][0][0])[q * 32 + lane] = 0.0;     // columns past m are never fetched
    for (int q = threadIdx.x; q < NW * 128; q += blockDim.x) xch[q] = 0.0;
    __syncthreads();
    for (int s0 = 0; s0 < n_strips; s0 += NW) {
        const int strip = s0 + warp;
        const int nwr = min(NW, n_strips - s0);
        const bool active = warp < nwr;
        const bool to_ring = warp + 1 < nwr;
        const int c0 = strip * DPSTRIP + lane * DPC;                 // 0-based first owned column
        double P0[DPC], P1[DPC];
#pragma unroll
        for (int c = 0; c < DPC; ++c) { P0[c] = MINF - open; P1[c] = 0.0; }      // row 0: (MIN - open, 0, 0)
        double out1 = 0.0, out2 = 0.0;
        double dsave = 0.0;
        const bool last_strip = strip == n_strips - 1;
        const double *xin = xch + (size_t)(warp > 0 ? warp - 1 : 0) * 128;
        double *xout = xch + (size_t)warp * 128;
        auto prefetch = [&](int t) {         // one (possibly empty) cp.async group per wavefront step
            const int i = t - lane + 1;
            if (i >= 1 && i <= n) {
                double (*slot)[32] = ring[t & (DTW_PF - 1)];
                const double *src = S + (long long)(i - 1) * m + c0;
#pragma unroll
                for (int c = 0; c < DPC; ++c)
                    if (c0 + c < m) dtw_cp_async8(&slot[c][lane], src + c);
                if (lane == 0 && warp == 0 && strip > 0) { dtw_cp_async8(&slot[DPC][0], bnd1 + i - 1); dtw_cp_async8(&slot[DPC + 1][0], bnd2 + i - 1); }
            }
            dtw_cp_commit();
        };
        if (active)
            for (int u = 0; u < DTW_PF; ++u) prefetch(u);
        const int T = n + 31;
        const int T4 = (T + 3) & ~3;
        const int total = T4 + DTW_SKEW * (nwr - 1);
        for (int G0 = 0; G0 < total; G0 += 4) {
            const int t0 = G0 - DTW_SKEW * warp;
            if (active && t0 >= 0 && t0 < T) {
#pragma unroll 1
                for (int t = t0; t < min(t0 + 4, T); ++t) {
                    const int i = t - lane + 1;      // 1-based row
                    const bool valid = i >= 1 && i <= n;
                    dtw_cp_wait<DTW_PF - 1>();
                    double (*slot)[32] = ring[t & (DTW_PF - 1)];
                    double sc[DPC];
#pragma unroll
                    for (int c = 0; c < DPC; ++c) sc[c] = slot[c][lane];
                    double b1 = 0.0, b2 = 0.0;
                    if (lane == 0) {
                        if (warp > 0) { const int e = (max(i, 1) - 1) & 63; b1 = xin[2 * e]; b2 = xin[2 * e + 1]; }
                        else { b1 = slot[DPC][0]; b2 = slot[DPC + 1][0]; }
                    }
                    prefetch(t + DTW_PF);
                    double L1 = shfl_up_d(out1), L2 = shfl_up_d(out2);
                    if (lane == 0) {
                        if (strip == 0) { L1 = 0.0; L2 = MINF - open; }       // column 0: (0, 0, MIN - open)
                        else if (valid) { L1 = b1; L2 = b2; }
                    }
                    if (i == 1) dsave = 0.0;          // M[0][j][1] = 0
                    const double in1 = L1;
                    double D1 = dsave;
                    if (valid) {
                        unsigned codes = 0;
#pragma unroll
                        for (int c = 0; c < DPC; ++c) {
                            const int j = c0 + c;     // 0-based column
                            const double s = sc[c];
                            const double l0 = P0[c] - ext, l1 = P1[c] - open;
                            const int ql = l1 > l0 ? 1 : 0;
                            const double lower = ql ? l1 : l0;
                            const double u0 = L1 - open, u1 = L2 - ext;
                            const int qu = u1 > u0 ? 1 : 0;
                            const double upper = qu ? u1 : u0;
                            const double dg = D1 + s;
                            double v = lower; int q = 0;
                            if (dg > v) { v = dg; q = 1; }
                            if (upper > v) { v = upper; q = 2; }
                            codes |= (unsigned)(ql | (q << 1) | (qu << 3)) << (8 * c);
                            if (i == n && j == m - 1) { final3[blockIdx.x * 3] = lower; final3[blockIdx.x * 3 + 1] = v; final3[blockIdx.x * 3 + 2] = upper; }
                            D1 = P1[c];
                            P0[c] = lower; P1[c] = v;
                            L1 = v; L2 = upper;
                        }
                        if (c0 < m) *reinterpret_cast<unsigned *>(B + (long long)(i - 1) * mp + c0) = codes;
                        out1 = L1; out2 = L2;
                        dsave = in1;
                        if (!last_strip && lane == 31) {
                            if (to_ring) { const int e = (i - 1) & 63; xout[2 * e] = out1; xout[2 * e + 1] = out2; }
                            else { bnd1[i - 1] = out1; bnd2[i - 1] = out2; }
                        }
                    }
                }
            }
            __syncthreads();
        }
        dtw_cp_wait<0>();
        __syncthreads();
    }
}

// k_dtw_fill for a batch: strips on concurrent warps where the batch is small enough to be latency-bound (at most mw_max problems)
// and has multi-strip problems; CARETTA_B200_DTW_MW=0: always the one-warp kernel.
inline cudaError_t launch_dtw_fill(const DpProblem *probs, int n_probs, int max_m, const double *S_all, unsigned char *B_all, double *bnd_all,
                                   double *final3, double open, double ext, cudaStream_t st, int mw_max = 128)
{
    const bool mw_on = !(getenv("CARETTA_B200_DTW_MW") && atoi(getenv("CARETTA_B200_DTW_MW")) == 0);
    const int strips = (max_m + DPSTRIP - 1) / DPSTRIP;
    const int nw = strips < DTW_MW_MAX ? strips : DTW_MW_MAX;
    if (mw_on && nw > 1 && n_probs <= mw_max) {
        const size_t sm = (size_t)nw * (DTW_PF * (DPC + 2) * 32 + 128) * sizeof(double);
        if (sm > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(k_dtw_fill_mw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
            if (e != cudaSuccess) return e;
        }
        k_dtw_fill_mw<<<n_probs, 32 * nw, sm, st>>>(probs, n_probs, S_all, B_all, bnd_all, final3, open, ext, nw);
    } else
        k_dtw_fill<<<n_probs, 32, 0, st>>>(probs, n_probs, S_all, B_all, bnd_all, final3, open, ext);
    return cudaGetLastError();
}

__global__ void k_dtw_trace(const DpProblem *probs, int n_probs, const unsigned char *B_all, const double *final3,
                            int *aln1, int *aln2, int *aln_len, double *score)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_probs) return;
    const DpProblem pr = probs[p];
    const unsigned char *B = B_all + pr.b_off;
    const double f0 = final3[p * 3], f1 = final3[p * 3 + 1], f2 = final3[p * 3 + 2];
    int dir = 0; double best = f0;
    if (f1 > best) { best = f1; dir = 1; }
    if (f2 > best) { best = f2; dir = 2; }
    score[p] = best;
    int *a1 = aln1 + pr.aln_off, *a2 = aln2 + pr.aln_off;
    int n = pr.n, m = pr.m, k = 0;
    const int mm = dtw_pitch(pr.m);
    while (!(n == 0 && m == 0)) {
        if (m == 0) { --n; a1[k] = n; a2[k] = -1; ++k; }
        else if (n == 0) { --m; a1[k] = -1; a2[k] = m; ++k; }
        else {
            const unsigned b = B[(long long)(n - 1) * mm + (m - 1)];
            if (dir == 0) { dir = b & 1; --n; a1[k] = n; a2[k] = -1; ++k; }
            else if (dir == 1) {
                dir = (b >> 1) & 3;
                if (dir == 1) { --n; --m; a1[k] = n; a2[k] = m; ++k; }
            } else { dir = ((b >> 3) & 1) + 1; --m; a1[k] = -1; a2[k] = m; ++k; }
        }
    }
    for (int x = 0, y = k - 1; x < y; ++x, --y) {
        int t1 = a1[x]; a1[x] = a1[y]; a1[y] = t1;
        int t2 = a2[x]; a2[x] = a2[y]; a2[y] = t2;
    }
    aln_len[p] = k;
}

// The same traceback with one warp per problem: the backtrack bytes the walk can reach next -- a tile of DTT_ROWS rows x DTT_COLS
// columns ending at the current cell -- are fetched by the whole warp (rows dealt to the lanes, aligned 32-bit loads) into shared
// memory, lane 0 walks inside the tile, and the warp reloads when the walk leaves it.  One memory round trip per tile instead
// of one per step; identical output.
constexpr int DTT_ROWS = 64, DTT_COLS = 128, DTT_WORDS = DTT_COLS / 4 + 1;      // (32 x 64 until the end of round 2: four times the reloads)

__global__ void __launch_bounds__(32) k_dtw_trace_w(const DpProblem *probs, int n_probs, const unsigned char *B_all, const double *final3,
                                                    int *aln1, int *aln2, int *aln_len, double *score)
{
    __shared__ unsigned tile[DTT_ROWS][DTT_WORDS + 1];
    __shared__ int tshift[DTT_ROWS];
    const int p = blockIdx.x, lane = threadIdx.x;
    if (p >= n_probs) return;
    const DpProblem pr = probs[p];
    const unsigned char *B = B_all + pr.b_off;
    const double f0 = final3[p * 3], f1 = final3[p * 3 + 1], f2 = final3[p * 3 + 2];
    int dir = 0; double best = f0;
    if (f1 > best) { best = f1; dir = 1; }
    if (f2 > best) { best = f2; dir = 2; }
    if (lane == 0) score[p] = best;
    int *a1 = aln1 + pr.aln_off, *a2 = aln2 + pr.aln_off;
    int n = pr.n, m = pr.m, k = 0;
    const int mm = dtw_pitch(pr.m);
    while (!(n == 0 && m == 0)) {
        // tile = rows r_lo..n-1, columns c_lo..m-1 of B (empty when the walk is on a border: no byte is needed there)
        const int r_hi = n - 1, c_hi = m - 1;
        const int r_lo = max(0, r_hi - DTT_ROWS + 1), c_lo = max(0, c_hi - DTT_COLS + 1);
        if (n > 0 && m > 0) {
#pragma unroll
            for (int rr = lane; rr < DTT_ROWS; rr += 32) {
                const int r = r_lo + rr;
                if (r <= r_hi) {
                    const unsigned long long addr = (unsigned long long)(B + (long long)r * mm + c_lo);
                    const unsigned *src = reinterpret_cast<const unsigned *>(addr & ~3ull);
                    const int sh = (int)(addr & 3ull);
                    const int words = (sh + (c_hi - c_lo + 1) + 3) >> 2;
#pragma unroll 4
                    for (int w = 0; w < words; ++w) tile[rr][w] = src[w];
                    tshift[rr] = sh;
                }
            }
        }
        __syncwarp();
        if (lane == 0) {
            while (!(n == 0 && m == 0)) {
                if (m == 0) { --n; a1[k] = n; a2[k] = -1; ++k; }
                else if (n == 0) { --m; a1[k] = -1; a2[k] = m; ++k; }
                else {
                    if (n - 1 < r_lo || m - 1 < c_lo) break;                 // left the tile
                    const int rr = n - 1 - r_lo, bo = tshift[rr] + (m - 1 - c_lo);
                    const unsigned b = (tile[rr][bo >> 2] >> ((bo & 3) * 8)) & 0xffu;
                    if (dir == 0) { dir = b & 1; --n; a1[k] = n; a2[k] = -1; ++k; }
                    else if (dir == 1) {
                        dir = (b >> 1) & 3;
                        if (dir == 1) { --n; --m; a1[k] = n; a2[k] = m; ++k; }
                    } else { dir = ((b >> 3) & 1) + 1; --m; a1[k] = -1; a2[k] = m; ++k; }
                }
            }
        }
        n = __shfl_sync(FULL, n, 0); m = __shfl_sync(FULL, m, 0); dir = __shfl_sync(FULL, dir, 0); k = __shfl_sync(FULL, k, 0);
        __syncwarp();
    }
    __syncwarp();
    for (int x = lane; x < k / 2; x += 32) {                                 // the walk wrote the path backwards
        const int y = k - 1 - x;
        const int t1 = a1[x], t2 = a2[x];
        a1[x] = a1[y]; a2[x] = a2[y];
        a1[y] = t1; a2[y] = t2;
    }
    if (lane == 0) aln_len[p] = k;
}

// ------------------------------------------------------------------------------------------------------------
// Smith-Waterman with a linear gap (any value), H stored in fp64; first row-major maximum; literal traceback.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) k_sw_fill(const DpProblem *probs, int n_probs, const double *S_all, double *H_all,
                                                double *bnd_all, double *best_val, long long *best_idx, double gap)
{
    if ((int)blockIdx.x >= n_probs) return;
    const DpProblem pr = probs[blockIdx.x];
    const int lane = threadIdx.x, n = pr.n, m = pr.m;
    const double *S = S_all + pr.s_off;
    double *H = H_all + pr.b_off;           // [n][m], cell (i,j) 1-based at (i-1)*m + (j-1)
    double *bnd = bnd_all + pr.bnd_off * 2;
    const int n_strips = (m + DPSTRIP - 1) / DPSTRIP;
    double bv = 0.0; long long bi = -1;      // first row-major maximum seen by this lane (strictly greater wins, then lowest index)
    for (int strip = 0; strip < n_strips; ++strip) {
        const int c0 = strip * DPSTRIP + lane * DPC;
        double P[DPC];
#pragma unroll
        for (int c = 0; c < DPC; ++c) P[c] = 0.0;
        double out = 0.0, dsave = 0.0;
        const bool last_strip = strip == n_strips - 1;
        for (int t = 0; t < n + 31; ++t) {
            const int i = t - lane + 1;
            const bool valid = i >= 1 && i <= n;
            double L = shfl_up_d(out);
            if (lane == 0) { L = 0.0; if (strip > 0 && valid) L = bnd[i - 1]; }
            if (i == 1) dsave = 0.0;
            const double in = L;
            double Dg = dsave;
            if (valid) {
#pragma unroll
                for (int c = 0; c < DPC; ++c) {
                    const int j = c0 + c;
                    double h = 0.0;
                    if (j < m) {
                        const double dg = Dg + S[(long long)(i - 1) * m + j];
                        const double lf = L - gap, up = P[c] - gap;
                        if (dg > h) h = dg;
                        if (lf > h) h = lf;
                        if (up > h) h = up;
                        const long long idx = (long long)(i - 1) * m + j;
                        H[idx] = h;
                        if (h > bv || (h == bv && bi >= 0 && idx < bi)) { if (h > 0.0) { bv = h; bi = idx; } }
                    }
                    Dg = P[c]; P[c] = h; L = h;
                }
                out = L; dsave = in;
                if (!last_strip && lane == 31) bnd[i - 1] = out;
            }
        }
        __syncwarp();
    }
    // warp argmax: highest value, then lowest row-major index
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(FULL, bv, o);
        const long long oi = __shfl_xor_sync(FULL, bi, o);
        if (oi >= 0 && (bi < 0 || ov > bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
    }
    if (lane == 0) { best_val[blockIdx.x] = bi >= 0 ? bv : 0.0; best_idx[blockIdx.x] = bi; }
}

__global__ void k_sw_trace(const DpProblem *probs, int n_probs, const double *S_all, const double *H_all, const double *best_val,
                           const long long *best_idx, int *aln1, int *aln2, int *aln_len, double *score, int *status, double gap,
                           int want_paths)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_probs) return;
    const DpProblem pr = probs[p];
    const double *S = S_all + pr.s_off, *H = H_all + pr.b_off;
    const int m = pr.m;
    score[p] = best_val[p];
    int st = 0, k = 0;
    if (best_idx[p] < 0) st = 2;           // no cell > 0: the reference raises (dynamic_time_warping.py:250)
    else if (want_paths) {
        int *a1 = aln1 + pr.aln_off, *a2 = aln2 + pr.aln_off;
        int i = (int)(best_idx[p] / m) + 1, j = (int)(best_idx[p] % m) + 1;
        auto h = [&](int ii, int jj) -> double { return (ii == 0 || jj == 0) ? 0.0 : H[(long long)(ii - 1) * m + (jj - 1)]; };
        while (i > 0 && j > 0) {
            const double sc = h(i, j);
            if (sc == 0.0) break;
            else if (sc == h(i - 1, j - 1) + S[(long long)(i - 1) * m + (j - 1)]) { --i; --j; a1[k] = i; a2[k] = j; ++k; }
            else if (sc == h(i, j - 1) - gap) { --j; a1[k] = -1; a2[k] = j; ++k; }
            else if (sc == h(i - 1, j) - gap) { --i; a1[k] = i; a2[k] = -1; ++k; }
            else break;
        }
        for (int x = 0, y = k - 1; x < y; ++x, --y) {
            int t1 = a1[x]; a1[x] = a1[y]; a1[y] = t1;
            int t2 = a2[x]; a2[x] = a2[y]; a2[y] = t2;
        }
    }
    if (aln_len) aln_len[p] = k;
    if (status) status[p] = st;
}

// ------------------------------------------------------------------------------------------------------------
// make_rmsd_coverage_tm_matrix(superpose_first=False): one thread per pair i<j over the A alignment columns.
// ------------------------------------------------------------------------------------------------------------
// bitsT[w * N + p] bit b <=> aln[p][32 w + b] != -1 (k_aln_bits): a pair only visits the columns where both have a residue
// (mask_i & mask_j), in ascending column order like the reference's loop, instead of scanning all A columns twice.
__global__ void k_rmsd_cov_tm(const long long *aln, int N, long long A, const unsigned *bitsT, int W, const double *coords,
                              const long long *offsets, const double *centroid, double *rmsd, double *cov, double *tm, int *n_bad,
                              int superpose)
{
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long np = (long long)N * (N - 1) / 2;
    if (q >= np) return;
    // unrank q -> (i, j), i < j, row-major over the upper triangle
    int i = (int)((2.0 * N - 1.0 - sqrt((2.0 * N - 1.0) * (2.0 * N - 1.0) - 8.0 * (double)q)) / 2.0);
    while ((long long)i * (2 * N - i - 1) / 2 > q) --i;
    while ((long long)(i + 1) * (2 * N - i - 2) / 2 <= q) ++i;
    const int j = (int)(q - (long long)i * (2 * N - i - 1) / 2) + i + 1;
    const long long *ai = aln + (long long)i * A, *aj = aln + (long long)j * A;
    const double *X = coords + offsets[i] * 3, *Y = coords + offsets[j] * 3;
    const double *ci = centroid + (long long)i * 3, *cj = centroid + (long long)j * 3;
    double s1[3] = {0, 0, 0}, s2[3] = {0, 0, 0}, Cr[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    int c = 0;
    for (int w = 0; w < W; ++w)
    for (unsigned mbits = bitsT[(long long)w * N + i] & bitsT[(long long)w * N + j]; mbits; mbits &= mbits - 1) {
        const long long k = (long long)w * 32 + (__ffs(mbits) - 1);
        const long long a = ai[k], b = aj[k];
        ++c;
        double x1[3], x2[3];
        for (int d = 0; d < 3; ++d) { x1[d] = X[a * 3 + d] - ci[d]; x2[d] = Y[b * 3 + d] - cj[d]; s1[d] += x1[d]; s2[d] += x2[d]; }
        for (int u = 0; u < 3; ++u)
            for (int v = 0; v < 3; ++v) Cr[u * 3 + v] += x2[u] * x1[v];
    }
    if (c < 3) { atomicAdd(n_bad, 1); return; }          // the reference asserts here; entries keep the diagonal defaults
    const double inv = 1.0 / (double)c;
    double p1[3], p2[3], m1[3], m2[3], Cm[9], R[9], tr[3];
    for (int d = 0; d < 3; ++d) { p1[d] = s1[d] * inv; p2[d] = s2[d] * inv; m1[d] = p1[d] + ci[d]; m2[d] = p2[d] + cj[d]; }
    for (int u = 0; u < 3; ++u)
        for (int v = 0; v < 3; ++v) Cm[u * 3 + v] = Cr[u * 3 + v] - s2[u] * p1[v];
    if (superpose) {
        kabsch_rotation(Cm, R);
        for (int b = 0; b < 3; ++b) tr[b] = m1[b] - (m2[0] * R[b] + m2[1] * R[3 + b] + m2[2] * R[6 + b]);
    } else {                                   // superpose_first=True: the chains were superposed beforehand (:1025-1026, :1037)
        for (int q = 0; q < 9; ++q) R[q] = (q % 4 == 0) ? 1.0 : 0.0;
        tr[0] = tr[1] = tr[2] = 0.0;
    }
    const long long l1 = offsets[i + 1] - offsets[i], l2 = offsets[j + 1] - offsets[j];
    const double d1 = 1.24 * (double)(l1 - 15) / 3 - 1.8, d2 = 1.24 * (double)(l2 - 15) / 3 - 1.8;
    double ss = 0.0, t1 = 0.0, t2 = 0.0;
    for (int w = 0; w < W; ++w)
    for (unsigned mbits = bitsT[(long long)w * N + i] & bitsT[(long long)w * N + j]; mbits; mbits &= mbits - 1) {
        const long long k = (long long)w * 32 + (__ffs(mbits) - 1);
        const long long a = ai[k], b = aj[k];
        const double y0 = Y[b * 3], y1 = Y[b * 3 + 1], y2 = Y[b * 3 + 2];
        double sm = 0.0;
        for (int d = 0; d < 3; ++d) {
            const double yr = superpose ? (y0 * R[d] + y1 * R[3 + d] + y2 * R[6 + d]) + tr[d] : (d == 0 ? y0 : (d == 1 ? y1 : y2));
            const double df = X[a * 3 + d] - yr;
            ss += df * df;
            sm += df;
        }
        const double q1 = sm / d1, q2 = sm / d2;
        t1 += 1 / (1 + q1 * q1);
        t2 += 1 / (1 + q2 * q2);
    }
    const double rr = sqrt(ss / (double)c);
    t1 = (1.0 / (double)l1) * t1;
    t2 = (1.0 / (double)l2) * t2;
    const double tt = t1 > t2 ? t1 : t2;
    const double cv = (double)c / (double)A;
    rmsd[(long long)i * N + j] = rmsd[(long long)j * N + i] = rr;
    cov[(long long)i * N + j] = cov[(long long)j * N + i] = cv;
    tm[(long long)i * N + j] = tm[(long long)j * N + i] = tt;
}

__global__ void k_fill_diag(double *rmsd, double *cov, double *tm, int N)
{
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= (long long)N * N) return;
    rmsd[q] = 0.0; cov[q] = 1.0; tm[q] = 1.0;          // multiple_alignment.py:1019-1024
}

}  // namespace crt
