// crt_fill_f32.cuh -- the fp32 production fills (the two hot kernels of the pair path), hand-scheduled for sm_100a.
//
// Same systolic structure as the generic k_fill in crt_kernels.cuh (one warp = one unit, lane = C columns, rows
// stream through the lanes), with the instruction count per DP cell cut down:
//   * the D-term dot product of the Gaussian exponent runs on packed FFMA2 (fma.rn.f32x2, sm_100): the record is
//     [r_0..r_{D-1}, A, 1] for a row and [r_0..r_{D-1}, 1, A] for a column, so A_a + B_b rides in the last pair and
//     the exponent costs (D+2)/2 FFMA2 + 1 FADD instead of D FFMA + 1 FADD.
//   * the difference-form recurrence keeps NEGATED horizontal differences: nu = a - d <= 0 and sd = S - d <= 0, so
//     the two traceback bits of a cell are the SIGN BITS of values that the recurrence needs anyway (one extra
//     FADD for sd), and a funnel shift pushes each sign bit into the code word: 3 instructions per cell for the
//     traceback instead of 5 (2 FSETP + 2 SEL + LEA).
//   * stage 2 evaluates two cells per instruction with FADD2/FMUL2/FFMA2 (columns stored negated, row broadcast once
//     per step).
//   * row records are read without index clamping (the arrays carry 32 records of padding on both sides) and are
//     prefetched into L1 eight rows ahead, which removes the exposed L2 latency of the lane that touches a new row.
#pragma once
#include "crt_kernels.cuh"

namespace crt {

constexpr int ROW_PAD = 48;          // records of padding before/after every row-record array
constexpr int PREFETCH_ROWS = 8;

__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// cp.async (LDGSTS): the warp stages the next 32 row records in shared memory one block ahead, so the per-step row
// reads are fixed-latency LDS and no lane ever waits on L2/HBM for a row nobody has touched yet.
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gsrc)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
constexpr int RING = 96;             // rows held in the shared-memory ring (3 blocks of 32)
constexpr int RING_OFF = 30;

struct Fill1Args {
    const float *rec;        // [ROW_PAD + sumL + ROW_PAD][RS] row/column records, pointer already past the front pad
    const int *meta;         // [ROW_PAD + sumL + ROW_PAD]
};

// ------------------------------------------------------------------------------------------------------------
// Stage 1: Smith-Waterman (gap 0) on the shape-tensor Gaussian, difference form, 2 traceback bits per cell.
// Word layout per lane and row: bit 2(C-1-c)+1 = (H != diag + S), bit 2(C-1-c) = (H != left) for owned column c.
//
// Software pipeline inside one basic block per step: the serial max/add chain of row t runs while the independent
// FFMA2 work of row t+1 is issued, and the record of row t+2 is in flight.
// ------------------------------------------------------------------------------------------------------------
// Shared-memory ring row: the RS floats of the record, the row meta (int bits) at [RS], padded to RS + 8 floats.
// The stride (80 B for d <= 10, 112 B for d <= 16) keeps the lanes' 128-bit LDS of consecutive rows on distinct banks.
template <int NP, int RS>
__device__ __forceinline__ void load_row1(const float *ring_row, float2 (&row)[NP], int &meta)
{
    const float4 *p4 = reinterpret_cast<const float4 *>(ring_row);
#pragma unroll
    for (int k = 0; k < NP / 2; ++k) {
        const float4 v = p4[k];
        row[2 * k] = make_float2(v.x, v.y);
        row[2 * k + 1] = make_float2(v.z, v.w);
    }
    if (NP & 1) row[NP - 1] = reinterpret_cast<const float2 *>(ring_row)[NP - 1];
    meta = __float_as_int(ring_row[RS]);
}

// Ring addressing: stream row g lives at ring row rr = g + RING_OFF, slot rr % 96.  Block B = ring rows [32B, 32B+32).
// With RING_OFF = 30 the loads of the window of steps [t0, t0+32) (row t+2-lane at step t) touch only blocks t0/32 and
// t0/32+1, so block t0/32+2 can be in flight during the whole window.
template <int RS>
__device__ __forceinline__ void stage_block1(float *srow, const float *rec_unit, const int *meta_unit, int B, int lane)
{
    constexpr int SROW = RS + 8;
    const int slot0 = (B % 3) * 32;
    const float *src = rec_unit + (long long)(32 * B - RING_OFF) * RS;  // record of stream row g = 32B - RING_OFF
    constexpr int CPR = RS / 4;                                         // 16-byte chunks per record
#pragma unroll
    for (int q = lane; q < 32 * CPR; q += 32) {
        const int row = q / CPR, part = q - row * CPR;
        cp_async16(srow + (slot0 + row) * SROW + part * 4, src + q * 4);
    }
    cp_async4(srow + (slot0 + lane) * SROW + RS, meta_unit + (32 * B - RING_OFF) + lane);
    cp_async_commit();
}

template <int NP, int C>
__device__ __forceinline__ void rbf_row1(const float2 (&row)[NP], const float2 (&col)[C][NP], float (&s)[C])
{
#pragma unroll
    for (int c = 0; c < C; ++c) {
        float2 e2 = __fmul2_rn(row[0], col[c][0]);
#pragma unroll
        for (int k = 1; k < NP; ++k) e2 = __ffma2_rn(row[k], col[c][k], e2);
        s[c] = ex2_approx(e2.x + e2.y);
    }
}

template <int D, int C, bool MULTI>
#ifndef CRT_FILL1_MINB
#define CRT_FILL1_MINB 1
#endif
__global__ void __launch_bounds__(32, CRT_FILL1_MINB) k_fill1_f32(const Unit *__restrict__ units, int n_units, Fill1Args args, FillOut out)
{
    constexpr int NP = (D + 2) / 2;                 // float2 pairs per record
    constexpr int RS = ((2 * NP + 3) / 4) * 4;      // record stride in floats
    static_assert(D % 2 == 0, "D must be even");
    const int lane = threadIdx.x;
    if ((int)blockIdx.x >= n_units) return;
    const Unit u = units[blockIdx.x];
    const int G = u.G;
    const int steps4 = u.tchunks * 4;
    float *bnd = MULTI ? reinterpret_cast<float *>(out.bnd) + u.bnd_base : nullptr;

    for (int strip = 0; strip < (MULTI ? u.n_strips : 1); ++strip) {
        float2 col[C][NP];
        const int c0 = (strip * 32 + lane) * C;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            if (c0 + c < u.m) {
                const float2 *p = reinterpret_cast<const float2 *>(args.rec + ((long long)u.col_base + c0 + c) * RS);
#pragma unroll
                for (int k = 0; k < NP; ++k) col[c][k] = p[k];
                const float2 t = col[c][NP - 1];
                col[c][NP - 1] = make_float2(t.y, t.x);            // (A, 1) -> (1, A)
            } else {
#pragma unroll
                for (int k = 0; k < NP; ++k) col[c][k] = make_float2(0.f, 0.f);
                col[c][NP - 1] = make_float2(0.f, -INFINITY);      // 2^-inf = 0: padded columns pass H[i][m] through
            }
        }
        const bool last_strip = !MULTI || strip == u.n_strips - 1;
        const bool emitter = last_strip && lane == 31;
        float nprev[C];                      // negated horizontal differences of the previous row: H[i-1][j-1] - H[i-1][j] <= 0
#pragma unroll
        for (int c = 0; c < C; ++c) nprev[c] = 0.f;
        float carry = 0.f;                   // v[i][cend] handed to lane+1
        float acc = 0.f;                     // running H[i][m] (last lane of the last strip)
        int istar = 0, r = 0;
        unsigned word = 0;
        uint4 *tbp = out.tb + u.tb_base + (long long)strip * u.tchunks * 32 + lane;
        const float *rec_unit = args.rec + u.row_base * RS;          // record of stream row g is rec_unit + g*RS
        const int *meta_unit = args.meta + u.row_base;
        constexpr int SROW = RS + 8;
        __shared__ __align__(16) float srow[RING * SROW];
        __syncwarp();
        stage_block1<RS>(srow, rec_unit, meta_unit, 0, lane);
        stage_block1<RS>(srow, rec_unit, meta_unit, 1, lane);
        cp_async_wait_all();
        __syncwarp();
        stage_block1<RS>(srow, rec_unit, meta_unit, 2, lane);
        // float offset in the ring of the row fetched at the end of step t (stream row t + 2 - lane): starts at slot 32 - lane
        int roff = (RING_OFF + 2 - lane) * SROW;

        float s_cur[C];
        float2 row_nxt[NP];
        int meta_cur, meta_nxt, meta_prev = 0;
        {
            float2 row0[NP];
            // rows -lane and 1-lane (pipeline fill; lane 31's row -31 is clamped to ring row 0, any record will do)
            load_row1<NP, RS>(srow + max(roff - 2 * SROW, 0), row0, meta_cur);
            rbf_row1<NP, C>(row0, col, s_cur);
            load_row1<NP, RS>(srow + roff - SROW, row_nxt, meta_nxt);
        }

        for (int t0 = 0; t0 < steps4; t0 += 4) {
            unsigned w[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int t = t0 + q;
                const int g = t - lane;
                if (q == 0 && (t0 & 31) == 0 && t0 > 0) {
                    // rows of block t0/32 + 1 must have landed; refill the slot whose rows nobody needs any more
                    cp_async_wait_all();
                    __syncwarp();
                    stage_block1<RS>(srow, rec_unit, meta_unit, (t0 >> 5) + 2, lane);
                }
                float a = __shfl_up_sync(FULL, carry, 1);
                if (lane == 0) {
                    a = 0.f;
                    if (MULTI && strip > 0) a = bnd[min(max(g, 0), G - 1)];
                }
                // rare events, one branch: the chain that ended on the previous row reports, a new chain resets
                if (((meta_prev & 2) | (meta_cur & 1)) != 0) {
                    if ((meta_prev & 2) && emitter && (unsigned)(g - 1) < (unsigned)G) {
                        const int pidx = u.pair_base + (meta_prev >> 2) - u.row_chain0;
                        out.pair_score[pidx] = (double)acc;
                        out.pair_istar[pidx] = istar;
                    }
                    if (meta_cur & 1) {              // first residue of a chain: H[0][*] = 0
#pragma unroll
                        for (int c = 0; c < C; ++c) nprev[c] = 0.f;
                        acc = 0.f; istar = 0; r = 0;
                        if (lane == 0 && strip == 0 && (unsigned)g < (unsigned)G)
                            out.pair_zflag[u.pair_base + (meta_cur >> 2) - u.row_chain0] = (s_cur[0] == 0.f) ? 1 : 0;
                    }
                }
                // serial chain of row t ...
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const float s = s_cur[c];
                    const float nb = nprev[c];                       // -(H[i-1][j] - H[i-1][j-1])
                    const float d = fmaxf(fmaxf(s, a), -nb);         // H[i][j] - H[i-1][j-1]
                    const float sd = s - d;                          // < 0  <=>  H != diag + S
                    const float nu = a - d;                          // < 0  <=>  H != left
                    word = __funnelshift_l(__float_as_uint(sd), word, 1);
                    word = __funnelshift_l(__float_as_uint(nu), word, 1);
                    nprev[c] = nu;
                    a = d + nb;                                      // H[i][j] - H[i-1][j] >= 0
                }
                // ... overlapped with the independent score work of row t+1
                rbf_row1<NP, C>(row_nxt, col, s_cur);
                carry = a;
                ++r;
                acc += a;
                if (a > 0.f) istar = r;              // last row whose H[i][m] exceeds H[i-1][m] (meaningful on lane 31)
                w[q] = word;
                if (MULTI && !last_strip && lane == 31 && (unsigned)g < (unsigned)G) bnd[g] = carry;
                meta_prev = meta_cur; meta_cur = meta_nxt;
                load_row1<NP, RS>(srow + roff, row_nxt, meta_nxt);      // row t + 2 - lane, consumed by the next step's rbf
                roff = (roff == (RING - 1) * SROW) ? 0 : roff + SROW;
            }
            tbp[(long long)(t0 >> 2) * 32] = make_uint4(w[0], w[1], w[2], w[3]);
        }
        // the last row of the unit may have ended a chain on the final step
        if ((meta_prev & 2) && emitter && (unsigned)(steps4 - 1 - lane) < (unsigned)G) {
            const int pidx = u.pair_base + (meta_prev >> 2) - u.row_chain0;
            out.pair_score[pidx] = (double)acc;
            out.pair_istar[pidx] = istar;
        }
        if (MULTI) __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------------------
// Stage 2: smith_waterman_score on the superposed CA coordinates, absolute form, two cells per packed instruction.
// rows: float4 (x, y, z, meta) already scaled by sqrt(gamma_c log2 e); cols: float4 per residue, same scale.
// ------------------------------------------------------------------------------------------------------------
struct Fill2Args {
    const float4 *rows;      // per-batch stage-2 row records, pointer already past the front pad
    const float4 *cols;      // per residue
};

template <int C, bool MULTI>
__global__ void __launch_bounds__(32) k_fill2_f32(const Unit *__restrict__ units, int n_units, Fill2Args args, FillOut out)
{
    static_assert(C % 2 == 0, "C must be even");
    constexpr int CP = C / 2;
    const int lane = threadIdx.x;
    if ((int)blockIdx.x >= n_units) return;
    const Unit u = units[blockIdx.x];
    const int G = u.G;
    const int steps4 = u.tchunks * 4;
    float *bnd = MULTI ? reinterpret_cast<float *>(out.bnd) + u.bnd_base : nullptr;

    for (int strip = 0; strip < (MULTI ? u.n_strips : 1); ++strip) {
        float2 nx[CP], ny[CP], nz[CP];       // NEGATED column coordinates, two columns per register pair
        const int c0 = (strip * 32 + lane) * C;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float4 v = make_float4(-1e18f, -1e18f, -1e18f, 0.f);      // padded column: distance^2 ~ 3e36 -> S = 0
            if (c0 + c < u.m) { v = args.cols[(long long)u.col_base + c0 + c]; v.x = -v.x; v.y = -v.y; v.z = -v.z; }
            if (c & 1) { nx[c / 2].y = v.x; ny[c / 2].y = v.y; nz[c / 2].y = v.z; }
            else { nx[c / 2].x = v.x; ny[c / 2].x = v.y; nz[c / 2].x = v.z; }
        }
        const bool last_strip = !MULTI || strip == u.n_strips - 1;
        float prev[C];
#pragma unroll
        for (int c = 0; c < C; ++c) prev[c] = 0.f;
        float carry = 0.f, dsave = 0.f;
        // rows staged through a shared-memory ring (see k_fill1_f32); here the load of step t is stream row t+1-lane,
        // ring row rr = g + 31, so a window of 32 steps touches blocks t0/32 and t0/32+1 only
        const float4 *rows_unit = args.rows + u.rows2_base;
        __shared__ float4 srow2[RING];
        auto stage2 = [&](int B) {
            cp_async16(&srow2[(B % 3) * 32 + lane], rows_unit + (32 * B - 31) + lane);
            cp_async_commit();
        };
        __syncwarp();
        stage2(0); stage2(1);
        cp_async_wait_all();
        __syncwarp();
        stage2(2);
        int slot = 31 - lane;                // ring slot of stream row -lane
        float4 rv_nxt = srow2[slot];
        slot += 1;
        int meta_prev = 0;
        const bool emitter = last_strip && lane == 31;

        for (int t = 0; t < steps4; ++t) {
            const int g = t - lane;
            if ((t & 31) == 0 && t > 0) {
                cp_async_wait_all();
                __syncwarp();
                stage2((t >> 5) + 2);
            }
            const float4 rv = rv_nxt;
            rv_nxt = srow2[slot];
            slot = (slot == RING - 1) ? 0 : slot + 1;
            const int meta = __float_as_int(rv.w);
            float left = __shfl_up_sync(FULL, carry, 1);
            if (lane == 0) {
                left = 0.f;
                if (MULTI && strip > 0) left = bnd[min(max(g, 0), G - 1)];
            }
            if (((meta_prev & 2) | (meta & 1)) != 0) {
                if ((meta_prev & 2) && emitter && (unsigned)(g - 1) < (unsigned)G)
                    out.pair_score[u.pair_base + (meta_prev >> 2) - u.row_chain0] = (double)carry;
                if (meta & 1) {
#pragma unroll
                    for (int c = 0; c < C; ++c) prev[c] = 0.f;
                    dsave = 0.f;
                }
            }
            const float in = left;
            float diag = dsave;
            const float2 rx = make_float2(rv.x, rv.x), ry = make_float2(rv.y, rv.y), rz = make_float2(rv.z, rv.z);
            float s[C];
#pragma unroll
            for (int cp = 0; cp < CP; ++cp) {
                const float2 dx = __fadd2_rn(rx, nx[cp]), dy = __fadd2_rn(ry, ny[cp]), dz = __fadd2_rn(rz, nz[cp]);
                float2 e2 = __fmul2_rn(dx, dx);
                e2 = __ffma2_rn(dy, dy, e2);
                e2 = __ffma2_rn(dz, dz, e2);
                s[2 * cp] = ex2_approx(-e2.x);
                s[2 * cp + 1] = ex2_approx(-e2.y);
            }
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const float up = prev[c];
                const float h = fmaxf(fmaxf(diag + s[c], left), up);
                prev[c] = h;
                diag = up; left = h;
            }
            carry = left;
            dsave = in;
            meta_prev = meta;
            if (MULTI && !last_strip && lane == 31 && (unsigned)g < (unsigned)G) bnd[g] = carry;
        }
        if ((meta_prev & 2) && emitter && (unsigned)(steps4 - 1 - lane) < (unsigned)G)
            out.pair_score[u.pair_base + (meta_prev >> 2) - u.row_chain0] = (double)carry;
        if (MULTI) __syncwarp();
    }
}

}  // namespace crt
