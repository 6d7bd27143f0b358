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


// ------------------------------------------------------------------------------------------------------------
// k_fill1_v3: same arithmetic as k_fill1_v2, leaner control.
//   * rows are processed in groups of four; a group takes the CHECKED body (chain-boundary test, emission, reset: the
//     v2 body) only when some lane of the warp can meet a chain boundary inside it, which a warp-uniform comparison
//     against the next boundary step decides.  Between boundaries (~9 of 10 groups at 300-residue chains) the FAST body
//     runs: no meta loads, no per-row branch.
//   * the row ring (96 rows = 3 blocks of 32, as in v2) carries 3 mirror rows behind it, so a group's four rows never
//     wrap and are read with immediate offsets from one base advanced once per group.
// ------------------------------------------------------------------------------------------------------------
constexpr int RING3 = RING;          // 96
constexpr int RING3_OFF = RING_OFF;  // ring row of stream row g is (g + 30) mod 96 (see stage_block1 for the block timing)

template <int RS>
__device__ __forceinline__ void stage_block3(float *srow, const float *rec_unit, const int *meta_unit, int B, int lane)
{
    constexpr int SROW = RS + 8;
    const int slot0 = (B % 3) * 32;
    const float *src = rec_unit + (long long)(32 * B - RING3_OFF) * RS;     // record of stream row g = 32B - RING3_OFF
    constexpr int CPR = RS / 4;
#pragma unroll
    for (int q = lane; q < 32 * CPR; q += 32) {
        const int row = q / CPR, part = q - row * CPR;
        cp_async16(srow + (slot0 + row) * SROW + part * 4, src + q * 4);
    }
    cp_async4(srow + (slot0 + lane) * SROW + RS, meta_unit + (32 * B - RING3_OFF) + lane);
    if (slot0 == 0) {                                                       // mirror rows 0..2 behind the ring
        if (lane < 3 * CPR) {
            const int row = lane / CPR, part = lane - row * CPR;
            cp_async16(srow + (RING3 + row) * SROW + part * 4, src + lane * 4);
        }
        if (lane < 3) cp_async4(srow + (RING3 + lane) * SROW + RS, meta_unit + (32 * B - RING3_OFF) + lane);
    }
    cp_async_commit();
}

template <int D, int C, bool MULTI>
__global__ void __launch_bounds__(32, 1) k_fill1_v3(const Unit *__restrict__ units, int n_units, Fill1Args args, FillOut out,
                                                    const long long *__restrict__ offsets)
{
    constexpr int CP = C / 2;
    constexpr int RS = ((D + 2 + 3) / 4) * 4;
    constexpr int SROW = RS + 8;
    static_assert(C % 2 == 0, "C must be even");
    const int lane = threadIdx.x;
    if ((int)blockIdx.x >= n_units) return;
    const Unit u = units[blockIdx.x];
    const int G = u.G;
    const int steps4 = u.tchunks * 4;
    float *bnd = MULTI ? reinterpret_cast<float *>(out.bnd) + u.bnd_base : nullptr;
    __shared__ __align__(16) float srow[(RING3 + 3) * SROW];

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
        float nprev[C];
#pragma unroll
        for (int c = 0; c < C; ++c) nprev[c] = 0.f;
        float carry = 0.f, acc = 0.f;
        int istar_t = -1, start_t = 0;       // step of the last growth of H[i][m] / step of the chain's first row (this lane)
        unsigned word = 0;
        uint4 *tbp = out.tb + u.tb_base + (long long)strip * u.tchunks * 32 + lane;
        const float *rec_unit = args.rec + u.row_base * RS;
        const int *meta_unit = args.meta + u.row_base;
        __syncwarp();
        stage_block3<RS>(srow, rec_unit, meta_unit, 0, lane);
        stage_block3<RS>(srow, rec_unit, meta_unit, 1, lane);
        cp_async_wait_all();
        __syncwarp();
        stage_block3<RS>(srow, rec_unit, meta_unit, 2, lane);

        // chain boundaries of the unit's row stream, as steps of lane 0: B_k = offsets[row_chain0 + k] - row_base.
        // Lane l meets boundary B at step B + l; the row before it (last of its chain) at step B - 1 + l.
        int kb = 0;
        int nextB = 0;                        // boundary whose window [nextB - 1, nextB + 31] is not yet behind t0

        float s_cur[C];
        float row_nxt[D + 1];
        int meta_cur = 0, meta_nxt = 0, meta_prev = 0;
        {
            float row0[D + 1];
            // rows -lane and 1-lane (pipeline fill; lane 31's row -31 is clamped to ring row 0, any record will do)
            load_row_v2<D, RS>(srow + max(RING3_OFF - lane, 0) * SROW, row0, meta_cur);
            rbf_row_v2<D, CP>(row0, colp, bp, s_cur);
            load_row_v2<D, RS>(srow + (RING3_OFF + 1 - lane) * SROW, row_nxt, meta_nxt);
        }

        int roff = (RING3_OFF + 2 - lane) * SROW;          // float offset of stream row t0 + 2 - lane, t0 = 0
        // lane 0's inputs from the previous strip (bnd[g], g = t for lane 0): one 128-bit load per group, issued one
        // group ahead so the L2 latency never sits in front of the row's dependent chain (bnd is padded to steps4 + 4)
        float4 bq = make_float4(0.f, 0.f, 0.f, 0.f), bq_nxt = bq;
        if (MULTI && strip > 0 && lane == 0) bq_nxt = *reinterpret_cast<const float4 *>(bnd);

        for (int t0 = 0; t0 < steps4; t0 += 4) {
            if (MULTI && strip > 0) {
                bq = bq_nxt;
                if (lane == 0) bq_nxt = *reinterpret_cast<const float4 *>(bnd + t0 + 4);
            }
            if ((t0 & 31) == 0 && t0 > 0) {
                // rows of block t0/32 + 1 have landed; refill the block whose rows nobody needs any more
                cp_async_wait_all();
                __syncwarp();
                stage_block3<RS>(srow, rec_unit, meta_unit, (t0 >> 5) + 2, lane);
            }
            // ring base of the row fetched at the end of step t0 (stream row t0 + 2 - lane); the group reads base + q rows
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
                            const int pidx = u.pair_base + (meta_prev >> 2) - u.row_chain0;
                            out.pair_score[pidx] = (double)acc;
                            out.pair_istar[pidx] = istar_t >= start_t ? istar_t - start_t + 1 : 0;
                        }
                        if (meta_cur & 1) {
#pragma unroll
                            for (int c = 0; c < C; ++c) nprev[c] = 0.f;
                            acc = 0.f; start_t = t; istar_t = t - 1;
                            if (lane == 0 && strip == 0 && (unsigned)g < (unsigned)G)
                                out.pair_zflag[u.pair_base + (meta_cur >> 2) - u.row_chain0] = (s_cur[0] == 0.f) ? 1 : 0;
                        }
                    }
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const float s = s_cur[c];
                        const float nb = nprev[c];
                        const float d = fmaxf(fmaxf(s, a), -nb);
                        const int sd = __float_as_int(s) - __float_as_int(d);
                        const float nu = a - d;
                        word = __funnelshift_l((unsigned)sd, word, 1);
                        word = __funnelshift_l(__float_as_uint(nu), word, 1);
                        nprev[c] = nu;
                        a = d + nb;
                    }
                    rbf_row_v2<D, CP>(row_nxt, colp, bp, s_cur);
                    carry = a;
                    acc += a;
                    if (a > 0.f) istar_t = t;
                    w[q] = word;
                    if (MULTI && !last_strip && lane == 31 && (unsigned)g < (unsigned)G) bnd[g] = carry;
                    meta_prev = meta_cur; meta_cur = meta_nxt;
                    load_row_v2<D, RS>(gb + q * SROW, row_nxt, meta_nxt);
                }
                // boundaries whose window is behind the next group
                while (kb <= u.n_pairs && nextB + 31 < t0 + 4) {
                    ++kb;
                    nextB = kb <= u.n_pairs ? (int)(offsets[u.row_chain0 + kb] - u.row_base) : 0x3fffffff;
                }
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int t = t0 + q;
                    float a = __shfl_up_sync(FULL, carry, 1);
                    if (lane == 0) a = (MULTI && strip > 0) ? (q == 0 ? bq.x : q == 1 ? bq.y : q == 2 ? bq.z : bq.w) : 0.f;
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const float s = s_cur[c];
                        const float nb = nprev[c];
                        const float d = fmaxf(fmaxf(s, a), -nb);
                        const int sd = __float_as_int(s) - __float_as_int(d);
                        const float nu = a - d;
                        word = __funnelshift_l((unsigned)sd, word, 1);
                        word = __funnelshift_l(__float_as_uint(nu), word, 1);
                        nprev[c] = nu;
                        a = d + nb;
                    }
                    rbf_row_v2<D, CP>(row_nxt, colp, bp, s_cur);
                    carry = a;
                    acc += a;
                    if (a > 0.f) istar_t = t;
                    w[q] = word;
                    if (MULTI && !last_strip && lane == 31 && (unsigned)(t - 31) < (unsigned)G) bnd[t - 31] = carry;
                    // row t + 2 - lane; its meta is only needed when the NEXT group is a checked one
                    {
                        const float4 *p4 = reinterpret_cast<const float4 *>(gb + q * SROW);
                        float tmp[((D + 1 + 3) / 4) * 4];
#pragma unroll
                        for (int k = 0; k < (D + 1 + 3) / 4; ++k) {
                            const float4 v = p4[k];
                            tmp[4 * k] = v.x; tmp[4 * k + 1] = v.y; tmp[4 * k + 2] = v.z; tmp[4 * k + 3] = v.w;
                        }
#pragma unroll
                        for (int k = 0; k <= D; ++k) row_nxt[k] = tmp[k];
                    }
                }
                // leaving the fast body: the metas the checked body keeps in registers (rows t0+4-lane, t0+5-lane); the
                // rows of this group carried no flags, so meta_prev = 0 is exact
                meta_prev = 0;
                meta_cur = __float_as_int((gb + 2 * SROW)[RS]);
                meta_nxt = __float_as_int((gb + 3 * SROW)[RS]);
            }
            tbp[(long long)(t0 >> 2) * 32] = make_uint4(w[0], w[1], w[2], w[3]);
        }
        if ((meta_prev & 2) && emitter && (unsigned)(steps4 - 1 - lane) < (unsigned)G) {
            const int pidx = u.pair_base + (meta_prev >> 2) - u.row_chain0;
            out.pair_score[pidx] = (double)acc;
            out.pair_istar[pidx] = istar_t >= start_t ? istar_t - start_t + 1 : 0;
        }
        if (MULTI) __syncwarp();
    }
}

}  // namespace crt
