// crt_fill2_v3.cuh -- stage-2 fp32 fill (smith_waterman_score on the superposed CA coordinates), second schedule.
// Same arithmetic as k_fill2_f32 (two cells per packed instruction, absolute-form recurrence), with the control of
// k_fill1_v3: scores of row t+1 are computed while the max chain of row t runs (software pipeline), rows are processed
// in groups of four that take the CHECKED body (chain-boundary test, emission, reset) only when some lane of the warp
// can meet a chain boundary inside the group, and the row ring carries 3 mirror rows so a group reads its rows with
// immediate offsets from one base.
#pragma once
#include "crt_fill_f32.cuh"

// registers: 80 per thread lets one stage-2 warp sit beside two stage-1 warps on every SM sub-partition
#ifndef CRT_FILL2_MINB
#define CRT_FILL2_MINB 24
#endif

namespace crt {

__device__ __forceinline__ void stage_block2(float4 *srow2, const float4 *rows_unit, int B, int lane)
{
    const int slot0 = (B % 3) * 32;
    const float4 *src = rows_unit + (32 * B - RING_OFF);          // record of stream row g = 32B - RING_OFF
    cp_async16(&srow2[slot0 + lane], src + lane);
    if (slot0 == 0 && lane < 3) cp_async16(&srow2[RING + lane], src + lane);      // mirror rows
    cp_async_commit();
}

// Scores of one row against the lane's C columns: 2^(-|r - c|^2) on the scaled coordinates (score_functions.py:6-11 with
// gamma folded into the scale).  CRT_FILL2_DIRECT: the differences squared (6 packed instructions per two cells); default: the
// expanded form  -|r|^2 - |c|^2 + r . (2c)  with the column constant as the addend of the first FMA (4 packed instructions per
// two cells: the stage-2 fill is issue-bound, 9.6 -> 7.6 issue cycles per cell against 7.8 of MUFU.EX2).  The expanded form
// rounds at the magnitude of |c|^2 (scaled units: (0.208 A^-1 x)^2, ~20 for a 300-residue chain) instead of |r - c|^2: the
// exponent of a matched cell carries ~2e-6 of absolute error, random per cell, the pair score ~1e-6 relative (tolerance 1e-4).
template <int CP>
__device__ __forceinline__ void rbf_row2(const float4 rv, const float2 (&nx)[CP], const float2 (&ny)[CP], const float2 (&nz)[CP], const float2 (&bc)[CP],
                                         float (&s)[2 * CP])
{
    const float2 rx = make_float2(rv.x, rv.x), ry = make_float2(rv.y, rv.y), rz = make_float2(rv.z, rv.z);
#ifdef CRT_FILL2_DIRECT
#pragma unroll
    for (int cp = 0; cp < CP; ++cp) {
        const float2 dx = __fadd2_rn(rx, nx[cp]), dy = __fadd2_rn(ry, ny[cp]), dz = __fadd2_rn(rz, nz[cp]);
        float2 e2 = __fmul2_rn(dx, dx);
        e2 = __ffma2_rn(dy, dy, e2);
        e2 = __ffma2_rn(dz, dz, e2);
        s[2 * cp] = ex2_approx(-e2.x);
        s[2 * cp + 1] = ex2_approx(-e2.y);
    }
#else
    const float ar = -__fmaf_rn(rv.z, rv.z, __fmaf_rn(rv.y, rv.y, rv.x * rv.x));
    const float2 ar2 = make_float2(ar, ar);
#pragma unroll
    for (int cp = 0; cp < CP; ++cp) {
        float2 e2 = __ffma2_rn(rx, nx[cp], bc[cp]);
        e2 = __ffma2_rn(ry, ny[cp], e2);
        e2 = __ffma2_rn(rz, nz[cp], e2);
        e2 = __fadd2_rn(e2, ar2);
        s[2 * cp] = ex2_approx(e2.x);
        s[2 * cp + 1] = ex2_approx(e2.y);
    }
#endif
}

template <int C, bool MULTI>
__global__ void __launch_bounds__(32, MULTI ? 1 : CRT_FILL2_MINB) k_fill2_v3(const Unit *__restrict__ units, int n_units, Fill2Args args, FillOut out,
                                                 const long long *__restrict__ offsets)
{
    static_assert(C % 2 == 0, "C must be even");
    constexpr int CP = C / 2;
    const int lane = threadIdx.x;
    if ((int)blockIdx.x >= n_units) return;
    const Unit u = units[blockIdx.x];
    const int G = u.G;
    const int steps4 = u.tchunks * 4;
    float *bnd = MULTI ? reinterpret_cast<float *>(out.bnd) + u.bnd_base : nullptr;
    __shared__ float4 srow2[RING + 3];

    for (int strip = 0; strip < (MULTI ? u.n_strips : 1); ++strip) {
        // two columns per register pair.  CRT_FILL2_DIRECT: the NEGATED coordinates; default: TWICE the coordinates and -|c|^2
        float2 nx[CP], ny[CP], nz[CP], bc[CP];
        const int c0 = (strip * 32 + lane) * C;
#pragma unroll
        for (int c = 0; c < C; ++c) {
#ifdef CRT_FILL2_DIRECT
            float4 v = make_float4(-1e18f, -1e18f, -1e18f, 0.f);      // padded column: distance^2 ~ 3e36 -> S = 0
            if (c0 + c < u.m) { v = args.cols[(long long)u.col_base + c0 + c]; v.x = -v.x; v.y = -v.y; v.z = -v.z; }
            const float b = 0.f;
#else
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            float b = -INFINITY;                                      // padded column: 2^-inf = 0
            if (c0 + c < u.m) {
                v = args.cols[(long long)u.col_base + c0 + c];
                b = -__fmaf_rn(v.z, v.z, __fmaf_rn(v.y, v.y, v.x * v.x));
                v.x = 2.f * v.x; v.y = 2.f * v.y; v.z = 2.f * v.z;
            }
#endif
            if (c & 1) { nx[c / 2].y = v.x; ny[c / 2].y = v.y; nz[c / 2].y = v.z; bc[c / 2].y = b; }
            else { nx[c / 2].x = v.x; ny[c / 2].x = v.y; nz[c / 2].x = v.z; bc[c / 2].x = b; }
        }
        const bool last_strip = !MULTI || strip == u.n_strips - 1;
        const bool emitter = last_strip && lane == 31;
        float prev[C];
#pragma unroll
        for (int c = 0; c < C; ++c) prev[c] = 0.f;
        float carry = 0.f, dsave = 0.f;
        const float4 *rows_unit = args.rows + u.rows2_base;
        __syncwarp();
        stage_block2(srow2, rows_unit, 0, lane);
        stage_block2(srow2, rows_unit, 1, lane);
        cp_async_wait_all();
        __syncwarp();
        stage_block2(srow2, rows_unit, 2, lane);

        int kb = 0, nextB = 0;               // chain-boundary window, see k_fill1_v3
        int slot = RING_OFF + 2 - lane;      // ring row of stream row t0 + 2 - lane, t0 = 0
        float s_cur[C];
        float4 rv_nxt;
        int meta_cur, meta_nxt, meta_prev = 0;
        {
            const float4 rv0 = srow2[max(slot - 2, 0)];          // row -lane (lane 31: clamped, any record will do)
            meta_cur = __float_as_int(rv0.w);
            rbf_row2<CP>(rv0, nx, ny, nz, bc, s_cur);
            rv_nxt = srow2[slot - 1];                            // row 1 - lane
            meta_nxt = __float_as_int(rv_nxt.w);
        }

        // lane 0's inputs from the previous strip, one 128-bit load per group issued one group ahead (see k_fill1_v3)
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
                stage_block2(srow2, rows_unit, (t0 >> 5) + 2, lane);
            }
            const float4 *gb = srow2 + slot;
            slot += 4;
            if (slot >= RING) slot -= RING;
            const bool checked = t0 + 4 >= nextB;                // warp-uniform
            if (checked) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int g = t0 + q - lane;
                    float left = __shfl_up_sync(FULL, carry, 1);
                    if (lane == 0) left = (MULTI && strip > 0) ? (q == 0 ? bq.x : q == 1 ? bq.y : q == 2 ? bq.z : bq.w) : 0.f;
                    if (((meta_prev & 2) | (meta_cur & 1)) != 0) {
                        if ((meta_prev & 2) && emitter && (unsigned)(g - 1) < (unsigned)G) {
                            const int pidx = u.pair_base + (meta_prev >> 2) - u.row_chain0;
                            if (!out.skip_status || !(out.skip_status[pidx] & ST_TIE)) out.pair_score[pidx] = (double)carry;
                        }
                        if (meta_cur & 1) {
#pragma unroll
                            for (int c = 0; c < C; ++c) prev[c] = 0.f;
                            dsave = 0.f;
                        }
                    }
                    const float in = left;
                    float diag = dsave;
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const float up = prev[c];
                        const float h = fmaxf(fmaxf(diag + s_cur[c], left), up);
                        prev[c] = h;
                        diag = up; left = h;
                    }
                    rbf_row2<CP>(rv_nxt, nx, ny, nz, bc, s_cur);
                    carry = left;
                    dsave = in;
                    if (MULTI && !last_strip && lane == 31 && (unsigned)g < (unsigned)G) bnd[g] = carry;
                    meta_prev = meta_cur; meta_cur = meta_nxt;
                    rv_nxt = gb[q];
                    meta_nxt = __float_as_int(rv_nxt.w);
                }
                while (kb <= u.n_pairs && nextB + 31 < t0 + 4) {
                    ++kb;
                    nextB = kb <= u.n_pairs ? (int)(offsets[u.row_chain0 + kb] - u.row_base) : 0x3fffffff;
                }
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int t = t0 + q;
                    float left = __shfl_up_sync(FULL, carry, 1);
                    if (lane == 0) left = (MULTI && strip > 0) ? (q == 0 ? bq.x : q == 1 ? bq.y : q == 2 ? bq.z : bq.w) : 0.f;
                    const float in = left;
                    float diag = dsave;
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const float up = prev[c];
                        const float h = fmaxf(fmaxf(diag + s_cur[c], left), up);
                        prev[c] = h;
                        diag = up; left = h;
                    }
                    rbf_row2<CP>(rv_nxt, nx, ny, nz, bc, s_cur);
                    carry = left;
                    dsave = in;
                    if (MULTI && !last_strip && lane == 31 && (unsigned)(t - 31) < (unsigned)G) bnd[t - 31] = carry;
                    rv_nxt = gb[q];
                }
                // leaving the fast body: rows t0+3-lane (no flags), t0+4-lane, t0+5-lane
                meta_prev = 0;
                meta_cur = __float_as_int(gb[2].w);
                meta_nxt = __float_as_int(gb[3].w);
            }
        }
        if ((meta_prev & 2) && emitter && (unsigned)(steps4 - 1 - lane) < (unsigned)G) {
            const int pidx = u.pair_base + (meta_prev >> 2) - u.row_chain0;
            if (!out.skip_status || !(out.skip_status[pidx] & ST_TIE)) out.pair_score[pidx] = (double)carry;
        }
        if (MULTI) __syncwarp();
    }
}

}  // namespace crt
