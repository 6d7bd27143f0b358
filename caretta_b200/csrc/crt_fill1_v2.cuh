// crt_fill1_v2.cuh -- stage-1 fp32 fill, second instruction schedule (same recurrence, same outputs as k_fill1_f32).
//
// What changes against k_fill1_f32:
//   * the packed FMA pairs two COLUMNS of the lane instead of two feature dimensions: the row value is the scalar
//     broadcast operand of FFMA2 (SASS `FFMA2 Rd, Ra.F32, Rb.F32x2, Rc.F32x2`), the column pair (c_k[2p], c_k[2p+1])
//     lives in an aligned register pair, the accumulator starts at (B[2p], B[2p+1]) and a packed add brings in the
//     row's A.  The exponent of two cells costs D FFMA2 + 1 FADD2 and lands directly in two registers: no horizontal
//     add, no (A,1)/(1,A) pair -> 11 FMA-pipe cycles per cell instead of 13 at D = 10.
//   * the "H != diag + S" traceback bit is the sign of the INTEGER difference of the (non-negative) float bit
//     patterns of S and d, which moves one instruction per cell from the FMA pipe to the ALU pipe.
// Exponent operation order (k_trace's exact zero test follows it): e = B_b; e = fma(r_k, c_k, e) for k = 0..D-1; e += A_a.
#pragma once
#include "crt_fill_f32.cuh"

namespace crt {

template <int D, int RS>
__device__ __forceinline__ void load_row_v2(const float *ring_row, float (&row)[D + 1], int &meta)
{
    const float4 *p4 = reinterpret_cast<const float4 *>(ring_row);
    float tmp[((D + 1 + 3) / 4) * 4];
#pragma unroll
    for (int k = 0; k < (D + 1 + 3) / 4; ++k) {
        const float4 v = p4[k];
        tmp[4 * k] = v.x; tmp[4 * k + 1] = v.y; tmp[4 * k + 2] = v.z; tmp[4 * k + 3] = v.w;
    }
#pragma unroll
    for (int k = 0; k <= D; ++k) row[k] = tmp[k];
    meta = __float_as_int(ring_row[RS]);
}

template <int D, int CP>
__device__ __forceinline__ void rbf_row_v2(const float (&row)[D + 1], const float2 (&colp)[CP][D], const float2 (&bp)[CP], float (&s)[2 * CP])
{
#pragma unroll
    for (int p = 0; p < CP; ++p) {
#ifdef CRT_V2_SCALAR
        float ex = bp[p].x, ey = bp[p].y;
#pragma unroll
        for (int k = 0; k < D; ++k) { ex = __fmaf_rn(row[k], colp[p][k].x, ex); ey = __fmaf_rn(row[k], colp[p][k].y, ey); }
        s[2 * p] = ex2_approx(ex + row[D]);
        s[2 * p + 1] = ex2_approx(ey + row[D]);
#else
        float2 e = __ffma2_rn(make_float2(row[0], row[0]), colp[p][0], bp[p]);
#pragma unroll
        for (int k = 1; k < D; ++k) e = __ffma2_rn(make_float2(row[k], row[k]), colp[p][k], e);
        e = __fadd2_rn(e, make_float2(row[D], row[D]));
        s[2 * p] = ex2_approx(e.x);
        s[2 * p + 1] = ex2_approx(e.y);
#endif
    }
}

template <int D, int C, bool MULTI>
__global__ void __launch_bounds__(32, 1) k_fill1_v2(const Unit *__restrict__ units, int n_units, Fill1Args args, FillOut out)
{
    constexpr int CP = C / 2;
    constexpr int RS = ((D + 2 + 3) / 4) * 4;       // record stride in floats (layout of k_prep: r_0..r_{D-1}, A, 1, pad)
    static_assert(C % 2 == 0, "C must be even");
    const int lane = threadIdx.x;
    if ((int)blockIdx.x >= n_units) return;
    const Unit u = units[blockIdx.x];
    const int G = u.G;
    const int steps4 = u.tchunks * 4;
    float *bnd = MULTI ? reinterpret_cast<float *>(out.bnd) + u.bnd_base : nullptr;

    for (int strip = 0; strip < (MULTI ? u.n_strips : 1); ++strip) {
        float2 colp[CP][D], bp[CP];
        const int c0 = (strip * 32 + lane) * C;
#pragma unroll
        for (int p = 0; p < CP; ++p) {
            float v0[D + 1], v1[D + 1];
#pragma unroll
            for (int k = 0; k <= D; ++k) { v0[k] = 0.f; v1[k] = 0.f; }
            v0[D] = -INFINITY; v1[D] = -INFINITY;           // padded column: 2^-inf = 0, H[i][m] passes through
            if (c0 + 2 * p < u.m) {
                const float *q = args.rec + ((long long)u.col_base + c0 + 2 * p) * RS;
#pragma unroll
                for (int k = 0; k <= D; ++k) v0[k] = q[k];
            }
            if (c0 + 2 * p + 1 < u.m) {
                const float *q = args.rec + ((long long)u.col_base + c0 + 2 * p + 1) * RS;
#pragma unroll
                for (int k = 0; k <= D; ++k) v1[k] = q[k];
            }
#pragma unroll
            for (int k = 0; k < D; ++k) colp[p][k] = make_float2(v0[k], v1[k]);
            bp[p] = make_float2(v0[D], v1[D]);
        }
        const bool last_strip = !MULTI || strip == u.n_strips - 1;
        const bool emitter = last_strip && lane == 31;
        float nprev[C];                      // negated horizontal differences of the previous row (<= 0)
#pragma unroll
        for (int c = 0; c < C; ++c) nprev[c] = 0.f;
        float carry = 0.f, acc = 0.f;
        int istar = 0, r = 0;
        unsigned word = 0;
        uint4 *tbp = out.tb + u.tb_base + (long long)strip * u.tchunks * 32 + lane;
        const float *rec_unit = args.rec + u.row_base * RS;
        const int *meta_unit = args.meta + u.row_base;
        constexpr int SROW = RS + 8;
        __shared__ __align__(16) float srow[RING * SROW];
        __syncwarp();
        stage_block1<RS>(srow, rec_unit, meta_unit, 0, lane);
        stage_block1<RS>(srow, rec_unit, meta_unit, 1, lane);
        cp_async_wait_all();
        __syncwarp();
        stage_block1<RS>(srow, rec_unit, meta_unit, 2, lane);
        int roff = (RING_OFF + 2 - lane) * SROW;

        float s_cur[C];
        float row_nxt[D + 1];
        int meta_cur, meta_nxt, meta_prev = 0;
        {
            float row0[D + 1];
            load_row_v2<D, RS>(srow + max(roff - 2 * SROW, 0), row0, meta_cur);
            rbf_row_v2<D, CP>(row0, colp, bp, s_cur);
            load_row_v2<D, RS>(srow + roff - SROW, row_nxt, meta_nxt);
        }

        for (int t0 = 0; t0 < steps4; t0 += 4) {
            unsigned w[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int t = t0 + q;
                const int g = t - lane;
                if (q == 0 && (t0 & 31) == 0 && t0 > 0) {
                    cp_async_wait_all();
                    __syncwarp();
                    stage_block1<RS>(srow, rec_unit, meta_unit, (t0 >> 5) + 2, lane);
                }
                float a = __shfl_up_sync(FULL, carry, 1);
                if (lane == 0) {
                    a = 0.f;
                    if (MULTI && strip > 0) a = bnd[min(max(g, 0), G - 1)];
                }
                if (((meta_prev & 2) | (meta_cur & 1)) != 0) {
                    if ((meta_prev & 2) && emitter && (unsigned)(g - 1) < (unsigned)G) {
                        const int pidx = u.pair_base + (meta_prev >> 2) - u.row_chain0;
                        out.pair_score[pidx] = (double)acc;
                        out.pair_istar[pidx] = istar;
                    }
                    if (meta_cur & 1) {
#pragma unroll
                        for (int c = 0; c < C; ++c) nprev[c] = 0.f;
                        acc = 0.f; istar = 0; r = 0;
                        if (lane == 0 && strip == 0 && (unsigned)g < (unsigned)G)
                            out.pair_zflag[u.pair_base + (meta_cur >> 2) - u.row_chain0] = (s_cur[0] == 0.f) ? 1 : 0;
                    }
                }
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const float s = s_cur[c];
                    const float nb = nprev[c];
                    const float d = fmaxf(fmaxf(s, a), -nb);         // H[i][j] - H[i-1][j-1] >= +0
                    const int sd = __float_as_int(s) - __float_as_int(d);   // s, d >= +0: negative  <=>  s < d  <=>  H != diag + S
                    const float nu = a - d;                          // < 0  <=>  H != left
                    word = __funnelshift_l((unsigned)sd, word, 1);
                    word = __funnelshift_l(__float_as_uint(nu), word, 1);
                    nprev[c] = nu;
                    a = d + nb;
                }
                rbf_row_v2<D, CP>(row_nxt, colp, bp, s_cur);
                carry = a;
                ++r;
                acc += a;
                if (a > 0.f) istar = r;
                w[q] = word;
                if (MULTI && !last_strip && lane == 31 && (unsigned)g < (unsigned)G) bnd[g] = carry;
                meta_prev = meta_cur; meta_cur = meta_nxt;
                load_row_v2<D, RS>(srow + roff, row_nxt, meta_nxt);
                roff = (roff == (RING - 1) * SROW) ? 0 : roff + SROW;
            }
            tbp[(long long)(t0 >> 2) * 32] = make_uint4(w[0], w[1], w[2], w[3]);
        }
        if ((meta_prev & 2) && emitter && (unsigned)(steps4 - 1 - lane) < (unsigned)G) {
            const int pidx = u.pair_base + (meta_prev >> 2) - u.row_chain0;
            out.pair_score[pidx] = (double)acc;
            out.pair_istar[pidx] = istar;
        }
        if (MULTI) __syncwarp();
    }
}

}  // namespace crt
