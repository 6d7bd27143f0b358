// crt_fill1_v4.cuh -- stage-1 fp32 fill with TIE FLAGS: k_fill1_v3's schedule (two columns per packed FMA, fast / checked
// row groups, cp.async row ring) and a third bit per cell that says "the reference's float64 DP may decide this cell
// differently".  k_trace marks a pair whose walk meets such a cell (CRT_ST_TIE), and the host re-runs the marked pairs through
// the float64 parity kernels, so that the fp32 production mode takes the reference's path on every pair.
//
// Why a cell can differ (dynamic_time_warping.py:231-247, :260-277; measured with tools/tie_study.c on config C3):
//   * absorption (98 % of the differing pairs): the reference keeps H as float64, so increments below ulp(H)/2 vanish and
//     candidates that differ by less than ulp(H) tie EXACTLY; its equality traceback then takes diag > left > up.  The
//     difference form keeps those tiny increments and resolves the tie by them.  A cell is suspect when the winner is not
//     the diagonal and a candidate of HIGHER priority is within kappa * H of it (kappa = c * 2^-53).
//   * rounding (2 %): two candidates agree to within the fp32 noise of the scores (relative ~1e-6); suspect when ANY other
//     candidate is within eps * d of the winner (eps = 1e-4), whichever wins.
// One test covers both.  With round-down subtractions the three margins y = d - s, u = d - a, v = d - b are -0 exactly for the
// candidates that attain the maximum and positive for the strict losers, so
//     z  = min_u32(bits(y), bits(u), bits(v))      smallest margin among the strict losers (the -0s are the largest unsigned)
//     r  = (z - min(theta, y)) - eps * d            theta = kappa * H[i-1][last column of the lane]
// is negative exactly when a strict loser is within theta + eps d of the maximum while the diagonal loses (y > 0), or within
// eps d of it when the diagonal wins (y = -0 takes theta out).  The code bits are sign bits as before: sign(y) = (s attains the
// maximum), sign(u) = (a attains it), sign(r) = suspect; 3 bits per cell, 30 of the 32 bits of a lane's row word at C = 10.
// A lower-priority candidate within theta of a non-diagonal winner is flagged too (harmless: rare).
//
// Cost per cell against k_fill1_v3: y replaces the integer difference, u the negated difference, v is the value handed to
// the right as before; new are the 3-input unsigned minimum, one FMNMX, one FFMA (the threshold eps d + min(theta, y)), one integer
// subtraction (z - threshold on the bit patterns: its sign is the suspect bit, see CRT_V4_SUSPECT) and one funnel shift.
#pragma once
#include "crt_fill1_v2.cuh"

namespace crt {

struct TieArgs {
    float kappa;      // c * 2^-53: candidates closer than kappa * H tie in the reference's float64 H matrix
    float eps;        // relative closeness below which fp32 cannot order two candidates
};

__device__ __forceinline__ unsigned umin3(unsigned a, unsigned b, unsigned c) { return __vimin3_u32(a, b, c); }

// The suspect bit: sign of z - (eps d + min(theta, y)).  Default (round 2, last session): the threshold by one FFMA and the
// comparison as an INTEGER subtraction of the bit patterns (both are non-negative floats, or z = 0x80000000 when all three
// candidates attain the maximum: then the difference is positive) -- one instruction moves from the FMA pipe, which bounds this
// kernel (16.2 of its cycles per cell against 13 on the ALU pipe), to the ALU pipe.  CRT_V4_FSUB: the float form of the first
// version, z - min(theta, y) - eps d by FADD + FFMA.
#ifdef CRT_V4_FSUB
#define CRT_V4_SUSPECT()                                                                              \
        const unsigned rbits = __float_as_uint(__fmaf_rn(d, neg_eps, __uint_as_float(z) - fminf(th, y)));
#else
#define CRT_V4_SUSPECT()                                                                              \
        const unsigned rbits = z - __float_as_uint(__fmaf_rn(d, -neg_eps, fminf(th, y)));
#endif
// one row of the lane's C cells; a: value entering from the left (vertical difference of the left neighbour)
#define CRT_V4_ROW()                                                                                  \
    _Pragma("unroll")                                                                                 \
    for (int c = 0; c < C; ++c) {                                                                     \
        const float s = s_cur[c];                                                                     \
        const float b = uprev[c];                                                                     \
        const float d = fmaxf(fmaxf(s, a), b);                                                        \
        const float y = __fadd_rd(d, -s);                                                             \
        const float u = __fadd_rd(d, -a);                                                             \
        const float v = __fadd_rd(d, -b);                                                             \
        const unsigned z = umin3(__float_as_uint(y), __float_as_uint(u), __float_as_uint(v));         \
        CRT_V4_SUSPECT()                                                                              \
        word = __funnelshift_l(__float_as_uint(y), word, 1);                                          \
        word = __funnelshift_l(__float_as_uint(u), word, 1);                                          \
        word = __funnelshift_l(rbits, word, 1);                                                       \
        uprev[c] = u;                                                                                 \
        a = v;                                                                                        \
    }

#ifndef CRT_V4_MINB
#define CRT_V4_MINB 1
#endif
template <int D, int C, bool MULTI>
__global__ void __launch_bounds__(32, CRT_V4_MINB) k_fill1_v4(const Unit *__restrict__ units, int n_units, Fill1Args args, FillOut out,
                                                    const long long *__restrict__ offsets, TieArgs tie)
{
    constexpr int CP = C / 2;
    constexpr int RS = ((D + 2 + 3) / 4) * 4;
    constexpr int SROW = RS + 8;
    static_assert(C % 2 == 0 && 3 * C <= 32, "C must be even and 3 bits per cell must fit a 32-bit word");
    const int lane = threadIdx.x;
    if ((int)blockIdx.x >= n_units) return;
    const Unit u_ = units[blockIdx.x];
    const int G = u_.G;
    const int steps4 = u_.tchunks * 4;
    const float neg_eps = -tie.eps, kappa = tie.kappa;
    float *bnd = MULTI ? reinterpret_cast<float *>(out.bnd) + u_.bnd_base : nullptr;
    __shared__ __align__(16) float srow[(RING3 + 3) * SROW];

    for (int strip = 0; strip < (MULTI ? u_.n_strips : 1); ++strip) {
        float2 colp[CP][D], bp[CP];
        const int c0 = (strip * 32 + lane) * C;
#pragma unroll
        for (int p = 0; p < CP; ++p) {
            float v0[D + 1], v1[D + 1];
#pragma unroll
            for (int k = 0; k <= D; ++k) { v0[k] = 0.f; v1[k] = 0.f; }
            v0[D] = -INFINITY; v1[D] = -INFINITY;           // padded column: 2^-inf = 0, H[i][m] passes through
            if (c0 + 2 * p < u_.m) {
                const float *q = args.rec + ((long long)u_.col_base + c0 + 2 * p) * RS;
#pragma unroll
                for (int k = 0; k <= D; ++k) v0[k] = q[k];
            }
            if (c0 + 2 * p + 1 < u_.m) {
                const float *q = args.rec + ((long long)u_.col_base + c0 + 2 * p + 1) * RS;
#pragma unroll
                for (int k = 0; k <= D; ++k) v1[k] = q[k];
            }
#pragma unroll
            for (int k = 0; k < D; ++k) colp[p][k] = make_float2(v0[k], v1[k]);
            bp[p] = make_float2(v0[D], v1[D]);
        }
        const bool last_strip = !MULTI || strip == u_.n_strips - 1;
        const bool emitter = last_strip && lane == 31;
        float uprev[C];                      // horizontal differences of the previous row (>= 0, -0 where the left value won)
#pragma unroll
        for (int c = 0; c < C; ++c) uprev[c] = 0.f;
        float carry = 0.f, acc = 0.f;
        // steps of: the last growth of H[i][m]; the last growth above the float64 resolution; the chain's first row (this lane)
        int istar_t = -1, isig_t = -1, start_t = 0;
        unsigned word = 0;
        uint4 *tbp = out.tb + u_.tb_base + (long long)strip * u_.tchunks * 32 + lane;
        const float *rec_unit = args.rec + u_.row_base * RS;
        const int *meta_unit = args.meta + u_.row_base;
        __syncwarp();
        stage_block3<RS>(srow, rec_unit, meta_unit, 0, lane);
        stage_block3<RS>(srow, rec_unit, meta_unit, 1, lane);
        cp_async_wait_all();
        __syncwarp();
        stage_block3<RS>(srow, rec_unit, meta_unit, 2, lane);

        int kb = 0;
        int nextB = 0;                        // boundary whose window [nextB - 1, nextB + 31] is not yet behind t0 (see k_fill1_v3)

        float s_cur[C];
        float row_nxt[D + 1];
        int meta_cur = 0, meta_nxt = 0, meta_prev = 0;
        {
            float row0[D + 1];
            load_row_v2<D, RS>(srow + max(RING3_OFF - lane, 0) * SROW, row0, meta_cur);
            rbf_row_v2<D, CP>(row0, colp, bp, s_cur);
            load_row_v2<D, RS>(srow + (RING3_OFF + 1 - lane) * SROW, row_nxt, meta_nxt);
        }

        int roff = (RING3_OFF + 2 - lane) * SROW;
        float4 bq = make_float4(0.f, 0.f, 0.f, 0.f), bq_nxt = bq;
        if (MULTI && strip > 0 && lane == 0) bq_nxt = *reinterpret_cast<const float4 *>(bnd);

        for (int t0 = 0; t0 < steps4; t0 += 4) {
            if (MULTI && strip > 0) {
                bq = bq_nxt;
                if (lane == 0) bq_nxt = *reinterpret_cast<const float4 *>(bnd + t0 + 4);
            }
            if ((t0 & 31) == 0 && t0 > 0) {
                cp_async_wait_all();
                __syncwarp();
                stage_block3<RS>(srow, rec_unit, meta_unit, (t0 >> 5) + 2, lane);
            }
            const float *gb = srow + roff;
            roff += 4 * SROW;
            if (roff >= RING3 * SROW) roff -= RING3 * SROW;
            const bool checked = t0 + 4 >= nextB;            // warp-uniform
            unsigned w[4];
            if (checked) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int t = t0 + q;
                    const int g = t - lane;
                    float a = __shfl_up_sync(FULL, carry, 1);
                    if (lane == 0) a = (MULTI && strip > 0) ? (q == 0 ? bq.x : q == 1 ? bq.y : q == 2 ? bq.z : bq.w) : 0.f;
                    if (((meta_prev & 2) | (meta_cur & 1)) != 0) {
                        if ((meta_prev & 2) && emitter && (unsigned)(g - 1) < (unsigned)G) {
                            const int pidx = u_.pair_base + (meta_prev >> 2) - u_.row_chain0;
                            out.pair_score[pidx] = (double)acc;
                            out.pair_istar[pidx] = (istar_t >= start_t ? istar_t - start_t + 1 : 0) | (istar_t != isig_t ? ISTAR_TIE : 0);
                        }
                        if (meta_cur & 1) {
#pragma unroll
                            for (int c = 0; c < C; ++c) uprev[c] = 0.f;
                            acc = 0.f; start_t = t; istar_t = t - 1; isig_t = t - 1;
                            if (lane == 0 && strip == 0 && (unsigned)g < (unsigned)G)
                                out.pair_zflag[u_.pair_base + (meta_cur >> 2) - u_.row_chain0] = (s_cur[0] == 0.f) ? 1 : 0;
                        }
                    }
                    const float th = acc * kappa;
                    CRT_V4_ROW()
                    rbf_row_v2<D, CP>(row_nxt, colp, bp, s_cur);
                    carry = a;
                    acc += a;
                    if (a > 0.f) istar_t = t;
                    if (a > th) isig_t = t;
                    w[q] = word;
                    if (MULTI && !last_strip && lane == 31 && (unsigned)g < (unsigned)G) bnd[g] = carry;
                    meta_prev = meta_cur; meta_cur = meta_nxt;
                    load_row_v2<D, RS>(gb + q * SROW, row_nxt, meta_nxt);
                }
                while (kb <= u_.n_pairs && nextB + 31 < t0 + 4) {
                    ++kb;
                    nextB = kb <= u_.n_pairs ? (int)(offsets[u_.row_chain0 + kb] - u_.row_base) : 0x3fffffff;
                }
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int t = t0 + q;
                    float a = __shfl_up_sync(FULL, carry, 1);
                    if (lane == 0) a = (MULTI && strip > 0) ? (q == 0 ? bq.x : q == 1 ? bq.y : q == 2 ? bq.z : bq.w) : 0.f;
                    const float th = acc * kappa;
                    CRT_V4_ROW()
                    rbf_row_v2<D, CP>(row_nxt, colp, bp, s_cur);
                    carry = a;
                    acc += a;
                    if (a > 0.f) istar_t = t;
                    if (a > th) isig_t = t;
                    w[q] = word;
                    if (MULTI && !last_strip && lane == 31 && (unsigned)(t - 31) < (unsigned)G) bnd[t - 31] = carry;
                    {
                        const float4 *p4 = reinterpret_cast<const float4 *>(gb + q * SROW);
                        float tmp[((D + 1 + 3) / 4) * 4];
#pragma unroll
                        for (int k = 0; k < (D + 1 + 3) / 4; ++k) {
                            const float4 vv = p4[k];
                            tmp[4 * k] = vv.x; tmp[4 * k + 1] = vv.y; tmp[4 * k + 2] = vv.z; tmp[4 * k + 3] = vv.w;
                        }
#pragma unroll
                        for (int k = 0; k <= D; ++k) row_nxt[k] = tmp[k];
                    }
                }
                meta_prev = 0;
                meta_cur = __float_as_int((gb + 2 * SROW)[RS]);
                meta_nxt = __float_as_int((gb + 3 * SROW)[RS]);
            }
            tbp[(long long)(t0 >> 2) * 32] = make_uint4(w[0], w[1], w[2], w[3]);
        }
        if ((meta_prev & 2) && emitter && (unsigned)(steps4 - 1 - lane) < (unsigned)G) {
            const int pidx = u_.pair_base + (meta_prev >> 2) - u_.row_chain0;
            out.pair_score[pidx] = (double)acc;
            out.pair_istar[pidx] = (istar_t >= start_t ? istar_t - start_t + 1 : 0) | (istar_t != isig_t ? ISTAR_TIE : 0);
        }
        if (MULTI) __syncwarp();
    }
}

}  // namespace crt
