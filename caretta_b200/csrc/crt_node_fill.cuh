// crt_node_fill.cuh -- float64 stage-1 fill of the progressive-alignment nodes from PRECOMPUTED scores.
//
// A tree level runs stage 1 of score_function (multiple_alignment.py:321-349) for every node: smith_waterman on the tensor Gaussian
// of the node's two children, one pair per unit, one warp per unit.  In the generic parity kernel (k_fill<P1F64>) that warp
// evaluates the Gaussian itself -- ten sub / mul / add and a float64 exp per cell, ~70 FP64-pipe instructions against ~8 for
// the recurrence -- and near the root of the tree, where a level holds one to eight nodes, the level waits 0.6 ms for one warp.
// Here the scores of all nodes of the batch come from a cell-parallel kernel (k_pair_scores64: the same expression, the same
// operation order, the same exp, so the same bits), and the fill warp only runs the recurrence, with its four scores per step
// prefetched PF wavefront steps ahead by cp.async into a shared-memory ring (the scheme of k_dtw_fill, crt_dp_batch.cuh).
// Codes, start row, zero flag and H[n][m] are written exactly as k_fill<P1F64, C, false, true, MULTI> writes them (k_trace reads them).
#pragma once
#include "crt_kernels.cuh"

namespace crt {

struct NodeScoreArgs {
    const Unit *units;
    const double *rec;          // [sumL][D] raw tensors, zero padded (rec64 of k_prep)
    double neg_gamma;
    double *S;                  // scores of unit u at S + u.s_base, row-major [G][m]
    int D;
};

// grid (tiles, units): a block computes a tile of 16 rows x 64 columns (four rows per thread) from the 16 + 64 records it stages
// in shared memory; score_functions.py:11 in the reference's operation order (P1F64::score).
constexpr int NS_TR = 16, NS_TC = 64, NS_DMAX = 16;
__global__ void __launch_bounds__(256) k_pair_scores64(NodeScoreArgs a)
{
    const Unit u = a.units[blockIdx.y];
    const int tiles_c = (u.m + NS_TC - 1) / NS_TC, tiles_r = (u.G + NS_TR - 1) / NS_TR;
    if ((int)blockIdx.x >= tiles_c * tiles_r) return;
    const int tr = blockIdx.x / tiles_c, tc = blockIdx.x - tr * tiles_c;
    const int g0 = tr * NS_TR, c0 = tc * NS_TC;
    __shared__ double xs[NS_TR][NS_DMAX], ys[NS_TC][NS_DMAX + 1];
    const int D = a.D;
    for (int q = threadIdx.x; q < NS_TR * D; q += 256) {
        const int r = q / D, k = q - r * D;
        xs[r][k] = g0 + r < u.G ? a.rec[(u.row_base + g0 + r) * D + k] : 0.0;
    }
    for (int q = threadIdx.x; q < NS_TC * D; q += 256) {
        const int cc = q / D, k = q - cc * D;
        ys[cc][k] = c0 + cc < u.m ? a.rec[((long long)u.col_base + c0 + cc) * D + k] : 0.0;
    }
    __syncthreads();
    const int cc = threadIdx.x & (NS_TC - 1), rq = (threadIdx.x >> 6) * 4;
    if (c0 + cc >= u.m) return;
    double *S = a.S + u.s_base;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int g = g0 + rq + e;
        if (g >= u.G) break;
        double acc = 0.0;
        for (int k = 0; k < D; ++k) {
            const double t = __dsub_rn(xs[rq + e][k], ys[cc][k]);
            acc = __dadd_rn(acc, __dmul_rn(t, t));
        }
        S[(long long)g * u.m + c0 + cc] = exp(__dmul_rn(a.neg_gamma, acc));
    }
}

constexpr int NODE_PF = 8;      // wavefront steps of prefetch distance

__device__ __forceinline__ void node_cp_async8(void *smem_dst, const void *gmem_src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}

// One warp per unit (one pair per unit: G = the row chain's length).  Absolute form, equality codes, like the generic kernel.
template <int C, bool MULTI>
__global__ void __launch_bounds__(32) k_fill_s64(const Unit *__restrict__ units, int n_units, const double *__restrict__ S_all, FillOut out)
{
    const int lane = threadIdx.x;
    if ((int)blockIdx.x >= n_units) return;
    const Unit u = units[blockIdx.x];
    const int G = u.G, m = u.m;
    const int steps4 = u.tchunks * 4;           // >= G + 31
    const double *S = S_all + u.s_base;
    double *bnd = MULTI ? reinterpret_cast<double *>(out.bnd) + u.bnd_base : nullptr;
    __shared__ double ring[NODE_PF][C][32];
    const int pidx = u.pair_base;

    for (int strip = 0; strip < (MULTI ? u.n_strips : 1); ++strip) {
        const int c0 = (strip * 32 + lane) * C;
        const bool last_strip = !MULTI || strip == u.n_strips - 1;
        double prev[C];
#pragma unroll
        for (int c = 0; c < C; ++c) prev[c] = 0.0;
        double carry = 0.0, dsave = 0.0;
        int istar = 0, r = 0;
        uint4 *tbp = out.tb + u.tb_base + (long long)strip * u.tchunks * 32 + lane;
        // columns past m score 0 (the generic kernel's pad_col: exp(-inf)): their ring entries are zeroed once and never fetched
        __syncwarp();
        for (int q = 0; q < NODE_PF; ++q)
#pragma unroll
            for (int c = 0; c < C; ++c) ring[q][c][lane] = 0.0;
        __syncwarp();
        auto prefetch = [&](int t) {         // one (possibly empty) cp.async group per wavefront step
            const int g = t - lane;
            if ((unsigned)g < (unsigned)G) {
                const double *src = S + (long long)g * m + c0;
#pragma unroll
                for (int c = 0; c < C; ++c)
                    if (c0 + c < m) node_cp_async8(&ring[t & (NODE_PF - 1)][c][lane], src + c);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        for (int q = 0; q < NODE_PF; ++q) prefetch(q);

        for (int t0 = 0; t0 < steps4; t0 += 4) {
            unsigned w[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int t = t0 + q;
                const int g = t - lane;
                const bool valid = (unsigned)g < (unsigned)G;
                const int gc = min(max(g, 0), G - 1);
                asm volatile("cp.async.wait_group %0;" ::"n"(NODE_PF - 1) : "memory");      // the group of step t has landed
                double sc[C];
#pragma unroll
                for (int c = 0; c < C; ++c) sc[c] = valid ? ring[t & (NODE_PF - 1)][c][lane] : 0.0;
                prefetch(t + NODE_PF);               // refills this slot (every lane reads only its own entries)
                double in = shfl_up1(carry);
                if (lane == 0) {
                    in = 0.0;
                    if (MULTI && strip > 0) in = bnd[gc];
                }
                if (valid && g == 0) {               // first residue of the row chain: H[0][*] = 0
#pragma unroll
                    for (int c = 0; c < C; ++c) prev[c] = 0.0;
                    dsave = 0.0; istar = 0; r = 0;
                }
                unsigned word = 0;
                double left = in, diag = dsave;
                bool grew = false;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const double s = sc[c];
                    const double up = prev[c];
                    const double dg = diag + s;
                    const double h = max3<double>(dg, left, up);
                    word |= ((h != dg ? 2u : 0u) | (h != left ? 1u : 0u)) << (2 * (C - 1 - c));
                    if (c == C - 1) grew = h > up;
                    diag = up;
                    prev[c] = h;
                    left = h;
                }
                carry = left;
                dsave = in;
                ++r;
                if (grew) istar = r;                 // last row whose H[i][m] exceeds H[i-1][m] (meaningful on lane 31)
                w[q] = word;
                if (valid) {
                    if (MULTI && !last_strip && lane == 31) bnd[g] = carry;
                    if (lane == 0 && strip == 0 && g == 0) out.pair_zflag[pidx] = (sc[0] == 0.0) ? 1 : 0;
                    if (last_strip && lane == 31 && g == G - 1) {
                        out.pair_score[pidx] = carry;
                        out.pair_istar[pidx] = istar;
                    }
                }
            }
            tbp[(long long)(t0 >> 2) * 32] = make_uint4(w[0], w[1], w[2], w[3]);
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
    }
}


// ---------------------------------------------------------------------------------------------- strips on concurrent warps
// The float64 kernels hold four columns per lane, so a strip is 128 columns and a node of 300 x 300 is three strips -- run one
// after the other by the unit's single warp above (993 wavefront steps), twelve for the 1500-column nodes near the root.  Here a
// CTA gives every strip of a round its own warp: warp w runs strip s0 + w, NODE_SKEW steps behind warp w - 1, and the boundary
// column (H[g][last column of the strip]) goes from lane 31 of one warp to lane 0 of the next through a 64-entry ring in shared
// memory.  All warps step in lockstep, one __syncthreads per group of four steps (the group of a 128-bit code store): lane 31 of
// warp w writes row g at global step g + 31 + NODE_SKEW w, lane 0 of warp w + 1 reads it at g + NODE_SKEW (w + 1) -- five steps
// later, so always behind a barrier.  331 + 36 (strips - 1) steps instead of 331 strips.  Every cell sees the same operands in the same order as in
// k_fill_s64: codes, start row, zero flag and H[n][m] are bit-identical.  Rounds of NW strips follow one another when a unit has
// more strips than the CTA has warps (boundary of a round through the unit's global boundary column, as between the strips above).
constexpr int NODE_SKEW = 36;
constexpr int NODE_MW_MAX = 12;
template <int C>
__global__ void __launch_bounds__(32 * NODE_MW_MAX) k_fill_s64_mw(const Unit *__restrict__ units, int n_units, const double *__restrict__ S_all,
                                                                   FillOut out, int NW)
{
    extern __shared__ double node_dyn[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if ((int)blockIdx.x >= n_units) return;
    const Unit u = units[blockIdx.x];
    const int G = u.G, m = u.m;
    const int steps4 = u.tchunks * 4;           // >= G + 31
    const double *S = S_all + u.s_base;
    double *bnd = reinterpret_cast<double *>(out.bnd) + u.bnd_base;
    double (*ring)[C][32] = reinterpret_cast<double (*)[C][32]>(node_dyn + (size_t)warp * NODE_PF * C * 32);
    double *xch = node_dyn + (size_t)NW * NODE_PF * C * 32;          // [NW][64]: boundary column handed to the next warp
    const int pidx = u.pair_base;
    for (int q = threadIdx.x; q < NW * 64; q += blockDim.x) xch[q] = 0.0;

    for (int s0 = 0; s0 < u.n_strips; s0 += NW) {
        const int strip = s0 + warp;
        const int nwr = min(NW, u.n_strips - s0);           // warps with a strip in this round
        const bool active = warp < nwr;
        const int c0 = (strip * 32 + lane) * C;
        const bool last_strip = strip == u.n_strips - 1;
        const bool to_ring = warp + 1 < nwr;                 // the next strip runs in this round
        double prev[C];
#pragma unroll
        for (int c = 0; c < C; ++c) prev[c] = 0.0;
        double carry = 0.0, dsave = 0.0;
        int istar = 0, r = 0;
        uint4 *tbp = out.tb + u.tb_base + (long long)strip * u.tchunks * 32 + lane;
        for (int q = 0; q < NODE_PF; ++q)
#pragma unroll
            for (int c = 0; c < C; ++c) ring[q][c][lane] = 0.0;
        __syncthreads();
        auto prefetch = [&](int t) {         // one (possibly empty) cp.async group per wavefront step
            const int g = t - lane;
            if ((unsigned)g < (unsigned)G) {
                const double *src = S + (long long)g * m + c0;
#pragma unroll
                for (int c = 0; c < C; ++c)
                    if (c0 + c < m) node_cp_async8(&ring[t & (NODE_PF - 1)][c][lane], src + c);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        if (active)
            for (int q = 0; q < NODE_PF; ++q) prefetch(q);
        const double *xin = xch + (warp > 0 ? warp - 1 : 0) * 64;
        double *xout = xch + warp * 64;

        const int total = steps4 + NODE_SKEW * (nwr - 1);
        for (int T0 = 0; T0 < total; T0 += 4) {
            const int t0 = T0 - NODE_SKEW * warp;
            if (active && t0 >= 0 && t0 < steps4) {
                unsigned w[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int t = t0 + q;
                    const int g = t - lane;
                    const bool valid = (unsigned)g < (unsigned)G;
                    const int gc = min(max(g, 0), G - 1);
                    asm volatile("cp.async.wait_group %0;" ::"n"(NODE_PF - 1) : "memory");      // the group of step t has landed
                    double sc[C];
#pragma unroll
                    for (int c = 0; c < C; ++c) sc[c] = valid ? ring[t & (NODE_PF - 1)][c][lane] : 0.0;
                    prefetch(t + NODE_PF);               // refills this slot (every lane reads only its own entries)
                    double in = shfl_up1(carry);
                    if (lane == 0) {
                        in = 0.0;
                        if (warp > 0) in = xin[gc & 63];
                        else if (strip > 0) in = bnd[gc];
                    }
                    if (valid && g == 0) {               // first residue of the row chain: H[0][*] = 0
#pragma unroll
                        for (int c = 0; c < C; ++c) prev[c] = 0.0;
                        dsave = 0.0; istar = 0; r = 0;
                    }
                    unsigned word = 0;
                    double left = in, diag = dsave;
                    bool grew = false;
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const double s = sc[c];
                        const double up = prev[c];
                        const double dg = diag + s;
                        const double h = max3<double>(dg, left, up);
                        word |= ((h != dg ? 2u : 0u) | (h != left ? 1u : 0u)) << (2 * (C - 1 - c));
                        if (c == C - 1) grew = h > up;
                        diag = up;
                        prev[c] = h;
                        left = h;
                    }
                    carry = left;
                    dsave = in;
                    ++r;
                    if (grew) istar = r;                 // last row whose H[i][m] exceeds H[i-1][m] (meaningful on lane 31)
                    w[q] = word;
                    if (valid) {
                        if (!last_strip && lane == 31) {
                            if (to_ring) xout[g & 63] = carry;
                            else bnd[g] = carry;
                        }
                        if (lane == 0 && strip == 0 && g == 0) out.pair_zflag[pidx] = (sc[0] == 0.0) ? 1 : 0;
                        if (last_strip && lane == 31 && g == G - 1) {
                            out.pair_score[pidx] = carry;
                            out.pair_istar[pidx] = istar;
                        }
                    }
                }
                tbp[(long long)(t0 >> 2) * 32] = make_uint4(w[0], w[1], w[2], w[3]);
            }
            __syncthreads();
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
    }
}

}  // namespace crt
