// crt_fill_tc.cuh -- the fp32 production fills on the 5th-generation tensor cores (sm_100a): PAIR PER LANE.
//
// Why: the Gaussian exponent of the stage-1 score (score_functions.py:6-11) is  e(a, b) = A_a + B_b + sum_k r_k(a) c_k(b),
// a rank-(d+2) contraction  E = P C^T  between the residues of two chains.  The systolic kernels (crt_fill1_v4.cuh) spend
// 11 of their ~24 issue cycles per cell on it.  Here tcgen05.mma (kind::tf32, fp32 accumulators in tensor memory) produces
// E, and the CUDA cores only do what is left: ex2, the difference-form recurrence, the traceback codes and the tie flags.
//
// Work decomposition (one CTA = one ROUND):
//   * a round is one column chain j (the reference's seq2, multiple_alignment.py:164-169) and up to 128 PARTNER row chains
//     i < j of similar length.  Thread l of the four DP warps owns pair (i_l, j) -- TMEM lane l, the M index of the MMA.
//   * a step is one row of every partner at once: the A operand [128 x K] holds the current row record of every lane (each
//     thread stages its own), the B operand [columns x K] the records of chain j, D[l, b] = e(row of lane l, column b).
//     The thread then walks ITS row from left to right: no shuffles, no wavefront, no pipeline fill; per-row overhead
//     (operand staging, barriers, bookkeeping) is paid once per ~150 cells.
//   * the previous DP row of a lane (horizontal differences, one float per column) lives in TENSOR MEMORY next to the
//     exponent tiles (tcgen05.ld / tcgen05.st, lane-private, 256 KB per SM): no shared-memory traffic for the DP state.
//   * columns are processed in strips of <= 160 (TMEM budget of a CTA: 96 columns of exponent tiles in a ring of three
//     32-column tiles + 160 columns of state = 256, so two CTAs share an SM and hide each other's latencies); the value
//     crossing a strip boundary goes through a [rows][128] float array in global memory (L2).
//   * tf32 has 11 significant bits, so every operand is split  x = hi + lo  (both tf32) and E = hi.hi + hi.lo + lo.hi + lo.lo
//     is the sum of four K-blocks in ONE accumulation (K = 4 x 12 = 48: six K = 8 instructions per tile -- the tensor pipe is
//     ~15 % busy); A_a and B_b enter as three-level splits against ones, so they are exact.  What is left is the fp32
//     accumulation inside the tensor core.
//
// Traceback codes: 3 bits per cell as in k_fill1_v4 (S attains the maximum, left attains it, suspect), 96 bits per 32-column
// tile, one 16-byte store per tile: [pair][strip][row][tile] uint4 (x, y, z = the 96 bits in push order, MSB first).
#pragma once
#include "crt_fill1_v4.cuh"

namespace crt {

constexpr int TC_LANES = 128;            // partners per round = TMEM lanes = M of the MMA
constexpr int TC_TILE = 32;              // columns per exponent tile (N of one tcgen05.mma)
constexpr int TC_RING = 3;               // exponent tiles in flight
constexpr int TC_SC = 160;               // widest strip: state columns in TMEM
constexpr int TC_STATE_COL = TC_RING * TC_TILE;
constexpr int TC_TMEM_COLS = 256;
constexpr int TC_THREADS = 160;          // 4 DP warps + 1 MMA warp

struct TcRound {
    long long bnd_base;      // strip boundary values [max_rows][128] floats (n_strips > 1)
    int col_base;            // packed residue index of chain j
    int col_chain;           // j
    int m;                   // columns
    int n_strips, strip_w;   // strips of strip_w columns (multiple of 16, <= TC_SC); the last strip may be narrower
    int part_base, n_part;   // partners [part_base, part_base + n_part) of the batch's partner array
    int max_rows;            // steps per strip = longest partner
};

struct TcPartner {           // one lane of a round = one pair
    long long tb_base;       // uint4 index of the pair's traceback codes [strip][row][tile]
    long long rows2_base;    // stage-2 row records of the pair (float4 index, batch-local)
    long long path_base;     // path buffer (short2 index)
    int row_base;            // packed residue index of row chain i
    int row_chain;           // i
    int n;                   // rows
    int slot;                // result slot of the pair
    int round;               // its round (batch-local)
    int pad_;
};

namespace tcg {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}\n" ::"r"(bar), "r"(parity) : "memory");
}
// the same with a suspend-time hint: the thread sleeps in hardware until the phase completes (no issue slots burnt by the spin)
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
                 "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}\n" ::"r"(bar), "r"(parity), "r"(1000000u) : "memory");
}
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// canonical K-major operand layout without swizzle: core matrix = 8 rows x 16 bytes (128 contiguous bytes); the KCH core
// matrices of one 8-row group follow each other (leading byte offset 128), row groups follow at KCH * 128 bytes
template <int KCH>
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(128u >> 4) << 16;
    d |= (uint64_t)((KCH * 128u) >> 4) << 32;
    d |= (uint64_t)1 << 46;                                   // descriptor version of sm_100
    return d;
}
// instruction descriptor: D = fp32, A = B = tf32, both K-major, N, M
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
                 : "memory");
}
#define CRT_R16(v, o) "=r"(v[o]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]), "=r"(v[o + 6]), "=r"(v[o + 7]), \
                      "=r"(v[o + 8]), "=r"(v[o + 9]), "=r"(v[o + 10]), "=r"(v[o + 11]), "=r"(v[o + 12]), "=r"(v[o + 13]), "=r"(v[o + 14]), "=r"(v[o + 15])
#define CRT_W16(v, o) "r"(v[o]), "r"(v[o + 1]), "r"(v[o + 2]), "r"(v[o + 3]), "r"(v[o + 4]), "r"(v[o + 5]), "r"(v[o + 6]), "r"(v[o + 7]), \
                      "r"(v[o + 8]), "r"(v[o + 9]), "r"(v[o + 10]), "r"(v[o + 11]), "r"(v[o + 12]), "r"(v[o + 13]), "r"(v[o + 14]), "r"(v[o + 15])
// 32 lanes x 16 consecutive columns: thread t of the warp gets TMEM lane (lane base + t), registers = columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t *v)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                 : CRT_R16(v, 0) : "r"(taddr));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t *v)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n"
                 :: "r"(taddr), CRT_W16(v, 0) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float tf32_rn(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }

}  // namespace tcg

// float index of element (row r, k) of an operand with KCH 16-byte chunks per row
template <int KCH>
__device__ __forceinline__ int tc_op_index(int r, int k) { return ((r >> 3) * KCH + (k >> 2)) * 32 + (r & 7) * 4 + (k & 3); }

struct TcFill1Args {
    const float *rec;            // [sumL][RS] records r_0..r_{D-1}, A, 1 (k_prep), pointer past the front pad
    const TcRound *rounds;
    const TcPartner *partners;
    uint4 *tb;
    float *bnd;
    int *pair_istar;
    int *pair_zflag;
    double *pair_score;          // H[n][m] of the stage-1 Smith-Waterman
    int *counter;                // round counter of the launch (zeroed by the host): CTAs take rounds dynamically
    TieArgs tie;
    long long *prof;             // -DTC_PROFILE: cycles of DP warp 0 of CTA 0 spent waiting for tiles / in the row overhead / in total
};
#ifdef TC_PROFILE
#define TC_T(x) x = clock64()
#else
#define TC_T(x)
#endif

// Sixteen cells of the lane's row (half an exponent tile).  e: exponents from the tensor core, ub: horizontal differences of the
// previous row (in), of this row (out).  Same cell arithmetic as CRT_V4_ROW (crt_fill1_v4.cuh).  The 96 code bits of a tile
// are pushed MSB first into one shift register: HALF = 0 pushes bits 0..47 (w[0] complete, 16 bits stay in `word`), HALF = 1
// bits 48..95.  (Three separate words -- one per bit type -- measured 15 % slower: tools/cell_probe.cu V6.)
template <int HALF>
__device__ __forceinline__ void tc_cells16(const uint32_t *e, uint32_t *ub, float &a, float th, float neg_eps, unsigned &word, uint32_t (&w)[3])
{
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        constexpr int base = HALF * 48;
        const float s = ex2_approx(__uint_as_float(e[c]));
        const float b = __uint_as_float(ub[c]);
        const float d = fmaxf(fmaxf(s, a), b);
        const float y = __fadd_rd(d, -s);
        const float u = __fadd_rd(d, -a);
        const float v = __fadd_rd(d, -b);
        const unsigned z = umin3(__float_as_uint(y), __float_as_uint(u), __float_as_uint(v));
        const float r = __fmaf_rn(d, neg_eps, __uint_as_float(z) - fminf(th, y));
        word = __funnelshift_l(__float_as_uint(y), word, 1);
        if ((base + 3 * c + 1) % 32 == 0) w[(base + 3 * c + 1) / 32 - 1] = word;
        word = __funnelshift_l(__float_as_uint(u), word, 1);
        if ((base + 3 * c + 2) % 32 == 0) w[(base + 3 * c + 2) / 32 - 1] = word;
        word = __funnelshift_l(__float_as_uint(r), word, 1);
        if ((base + 3 * c + 3) % 32 == 0) w[(base + 3 * c + 3) / 32 - 1] = word;
        ub[c] = __float_as_uint(u);
        a = v;
    }
}

// RS: floats per record (12 for d <= 10, 20 for d <= 16).  K = 3 RS rounded up to a multiple of 8.
// Persistent CTAs: each takes rounds from g.counter until none is left (TMEM, barriers and their phases live across rounds).
template <int RS>
#ifdef TC_SPLIT3
__global__ void __launch_bounds__(TC_THREADS, 1) k_fill1_tc(TcFill1Args g, int n_rounds)
#else
__global__ void __launch_bounds__(TC_THREADS, 2) k_fill1_tc(TcFill1Args g, int n_rounds)
#endif
{
    constexpr int GCH = RS / 4;                          // 16-byte chunks per K-block (one record)
#ifdef TC_SPLIT3
    constexpr int K = ((6 * RS + 7) / 8) * 8;           // experiment: three-level split of the features, six product blocks
#else
    constexpr int K = ((4 * RS + 7) / 8) * 8;
#endif
    constexpr int KCH = K / 4;
    constexpr int D = RS == 12 ? 10 : 16;                // index of A inside a record
    constexpr int A_FLOATS = TC_LANES * K, B_FLOATS = TC_SC * K;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float *opA = reinterpret_cast<float *>(smem_raw);               // two buffers [128][K]
    float *opB = opA + 2 * A_FLOATS;                                // [TC_SC][K]
    __shared__ __align__(8) unsigned long long bars[2 + 2 * TC_RING];
    __shared__ uint32_t tmem_base_s;
    __shared__ int round_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar_a = tcg::smem_u32(&bars[0]);                 // a_full[2]
    const uint32_t bar_ef = tcg::smem_u32(&bars[2]);                // e_full[TC_RING]
    const uint32_t bar_ee = tcg::smem_u32(&bars[2 + TC_RING]);      // e_free[TC_RING]

    if (warp == 4) {
        if (lane == 0) {
            for (int q = 0; q < 2; ++q) tcg::mbar_init(bar_a + 8 * q, TC_LANES);
            for (int q = 0; q < TC_RING; ++q) { tcg::mbar_init(bar_ef + 8 * q, 1); tcg::mbar_init(bar_ee + 8 * q, 4); }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tcg::smem_u32(&tmem_base_s)), "n"(TC_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    // zero both A buffers and the B buffer once: the padding chunks beyond 3 RS stay zero
    for (int q = tid; q < (2 * A_FLOATS + B_FLOATS) / 4; q += TC_THREADS) reinterpret_cast<float4 *>(opA)[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    tcg::fence_async_smem();
    tcg::fence_before();
    __syncthreads();
    tcg::fence_after();
    const uint32_t tmem = tmem_base_s;
    const bool dp = warp < 4;
    const float neg_eps = -g.tie.eps, kappa = g.tie.kappa;
    const uint32_t lane_taddr = tmem + ((uint32_t)(warp * 32) << 16);
    uint32_t a_buf = 0, a_par = 0;          // A buffer of the current (strip, row) iteration and the parity of its barrier
    uint32_t e_slot = 0, e_par = 0;         // ring slot of the next exponent tile and the parity of its use

  for (;;) {
    __syncthreads();
    if (tid == 0) round_s = atomicAdd(g.counter, 1);
    __syncthreads();
    const int round = round_s;
    if (round >= n_rounds) break;
    const TcRound R = g.rounds[round];

    // ---- per-lane pair
    const bool have = dp && tid < R.n_part;
    TcPartner P{};
    if (have) P = g.partners[R.part_base + tid];
    const int n_l = have ? P.n : 1;
    const float *rec_l = g.rec + (long long)(have ? P.row_base : R.col_base) * RS;     // idle lanes walk chain j's first row

    // the lane's record of row r (rows past the end of a shorter partner repeat its last row: their cells are never read)
    float4 xr[GCH];
    auto fetch_row = [&](int r) {
        const float4 *src = reinterpret_cast<const float4 *>(rec_l + (long long)min(r, n_l - 1) * RS);
#pragma unroll
        for (int q = 0; q < GCH; ++q) xr[q] = __ldg(src + q);
    };
    // stage it into A buffer `buf`:  [hi(r), A_hi, 1 | hi(r), A_mid, 1 | lo(r), A_lo, 1 | lo(r), 0, 0]
    auto stage_row = [&](uint32_t buf) {
        float x[RS], hi[RS], lo[RS], mid[RS], l2[RS];
#pragma unroll
        for (int q = 0; q < GCH; ++q) { x[4 * q] = xr[q].x; x[4 * q + 1] = xr[q].y; x[4 * q + 2] = xr[q].z; x[4 * q + 3] = xr[q].w; }
#pragma unroll
        for (int k = 0; k < RS; ++k) { hi[k] = tcg::tf32_rn(x[k]); lo[k] = x[k] - hi[k]; mid[k] = hi[k]; l2[k] = lo[k]; }
        mid[D] = tcg::tf32_rn(lo[D]);                    // A = hi + mid + lo, against the column's ones
        lo[D] = lo[D] - mid[D];
        lo[D + 1] = 1.f;                                 // the ones that carry the column's B_mid, B_lo
        l2[D] = 0.f; l2[D + 1] = 0.f;                    // fourth block: lo(r) . lo(c) only
        float *dst = opA + buf * A_FLOATS + ((tid >> 3) * KCH) * 32 + (tid & 7) * 4;
#pragma unroll
        for (int q = 0; q < GCH; ++q) {
            *reinterpret_cast<float4 *>(dst + q * 32) = make_float4(hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
            *reinterpret_cast<float4 *>(dst + (GCH + q) * 32) = make_float4(mid[4 * q], mid[4 * q + 1], mid[4 * q + 2], mid[4 * q + 3]);
            *reinterpret_cast<float4 *>(dst + (2 * GCH + q) * 32) = make_float4(lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
            *reinterpret_cast<float4 *>(dst + (3 * GCH + q) * 32) = make_float4(l2[4 * q], l2[4 * q + 1], l2[4 * q + 2], l2[4 * q + 3]);
        }
#ifdef TC_SPLIT3
        {   // blocks: [h, A_h, 1 | h, A_m, 1 | m, A_l, 1 | m, 0, 0 | h, 0, 0 | l, 0, 0] with x = h + m + l exactly (tf32 each)
            float m3[RS], l3[RS], h0[RS], b2[RS];
#pragma unroll
            for (int k = 0; k < RS; ++k) { const float r1 = x[k] - hi[k]; m3[k] = tcg::tf32_rn(r1); l3[k] = r1 - m3[k]; h0[k] = hi[k]; b2[k] = m3[k]; }
            b2[D] = lo[D]; b2[D + 1] = 1.f;           // third block carries A_l and the one
            m3[D] = 0.f; m3[D + 1] = 0.f; h0[D] = 0.f; h0[D + 1] = 0.f; l3[D] = 0.f; l3[D + 1] = 0.f;
#pragma unroll
            for (int q = 0; q < GCH; ++q) {
                *reinterpret_cast<float4 *>(dst + (2 * GCH + q) * 32) = make_float4(b2[4 * q], b2[4 * q + 1], b2[4 * q + 2], b2[4 * q + 3]);
                *reinterpret_cast<float4 *>(dst + (3 * GCH + q) * 32) = make_float4(m3[4 * q], m3[4 * q + 1], m3[4 * q + 2], m3[4 * q + 3]);
                *reinterpret_cast<float4 *>(dst + (4 * GCH + q) * 32) = make_float4(h0[4 * q], h0[4 * q + 1], h0[4 * q + 2], h0[4 * q + 3]);
                *reinterpret_cast<float4 *>(dst + (5 * GCH + q) * 32) = make_float4(l3[4 * q], l3[4 * q + 1], l3[4 * q + 2], l3[4 * q + 3]);
            }
        }
#endif
        tcg::fence_async_smem();
        tcg::mbar_arrive(bar_a + 8 * buf);
    };

    for (int strip = 0; strip < R.n_strips; ++strip) {
        const int c0 = strip * R.strip_w;
        const int w_cols = min(R.strip_w, ((R.m - c0 + 15) / 16) * 16);      // columns of this strip, multiple of 16
        const int n_full = w_cols / TC_TILE;                                  // 32-column tiles
        const bool half_last = (w_cols & 16) != 0;                            // plus one 16-column tile
        const int n_tiles = n_full + (half_last ? 1 : 0);
        const bool last_strip = strip == R.n_strips - 1;
        // ---- column operand of the strip: [hi(c), 1, B_hi | lo(c), 1, B_mid | hi(c), 1, B_lo | lo(c), 0, 0]; padded columns: B_hi = -1e30
        if (strip > 0) __syncthreads();                                       // every MMA of the previous strip has been consumed
        for (int q = tid; q < n_tiles * TC_TILE * GCH; q += TC_THREADS) {
            const int c = q / GCH, part = q - c * GCH;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            const bool real = c0 + c < R.m;
            if (real) v = __ldg(reinterpret_cast<const float4 *>(g.rec + ((long long)R.col_base + c0 + c) * RS) + part);
            float x[4] = {v.x, v.y, v.z, v.w}, hi[4], lo[4], h2[4], l2[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) { hi[k] = tcg::tf32_rn(x[k]); lo[k] = x[k] - hi[k]; h2[k] = hi[k]; l2[k] = lo[k]; }
            if (part == D / 4) {                                              // the chunk that holds (A, 1): becomes (1, B)
                constexpr int ka = D % 4;                                     // position of A inside the chunk (2 for D = 10, 0 for D = 16)
                const float B = real ? x[ka] : -1e30f;
                const float bh = tcg::tf32_rn(B), r1 = B - bh, bm = tcg::tf32_rn(r1), bl = r1 - bm;
                hi[ka] = 1.f; hi[ka + 1] = bh;
                lo[ka] = 1.f; lo[ka + 1] = bm;
                h2[ka] = 1.f; h2[ka + 1] = bl;
                l2[ka] = 0.f; l2[ka + 1] = 0.f;
            }
            float *dst = opB + ((c >> 3) * KCH) * 32 + (c & 7) * 4;
            *reinterpret_cast<float4 *>(dst + part * 32) = make_float4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<float4 *>(dst + (GCH + part) * 32) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            *reinterpret_cast<float4 *>(dst + (2 * GCH + part) * 32) = make_float4(h2[0], h2[1], h2[2], h2[3]);
            *reinterpret_cast<float4 *>(dst + (3 * GCH + part) * 32) = make_float4(l2[0], l2[1], l2[2], l2[3]);
#ifdef TC_SPLIT3
            {   // blocks: [h, 1, B_h | m, 1, B_m | h, 1, B_l | m, 0, 0 | l, 0, 0 | h, 0, 0]
                float m3[4], l3[4], m1[4], h5[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) { const float r1 = x[k] - tcg::tf32_rn(x[k]); m3[k] = tcg::tf32_rn(r1); l3[k] = r1 - m3[k]; m1[k] = m3[k]; h5[k] = tcg::tf32_rn(x[k]); }
                if (part == D / 4) {
                    constexpr int ka = D % 4;
                    m1[ka] = lo[ka]; m1[ka + 1] = lo[ka + 1];      // (1, B_m) as computed above
                    m3[ka] = 0.f; m3[ka + 1] = 0.f; l3[ka] = 0.f; l3[ka + 1] = 0.f; h5[ka] = 0.f; h5[ka + 1] = 0.f;
                }
                *reinterpret_cast<float4 *>(dst + (GCH + part) * 32) = make_float4(m1[0], m1[1], m1[2], m1[3]);
                *reinterpret_cast<float4 *>(dst + (3 * GCH + part) * 32) = make_float4(m3[0], m3[1], m3[2], m3[3]);
                *reinterpret_cast<float4 *>(dst + (4 * GCH + part) * 32) = make_float4(l3[0], l3[1], l3[2], l3[3]);
                *reinterpret_cast<float4 *>(dst + (5 * GCH + part) * 32) = make_float4(h5[0], h5[1], h5[2], h5[3]);
            }
#endif
        }
        tcg::fence_async_smem();
        if (dp) {
            fetch_row(0);
            stage_row(a_buf);
            fetch_row(1);
            // H[0][*] = 0: the state of the strip starts as zeros
            uint32_t zero[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) zero[c] = 0u;
            for (int c = 0; c < w_cols; c += 16) tcg::tmem_st16(lane_taddr + TC_STATE_COL + c, zero);
            tcg::tmem_wait_st();
        }
        __syncthreads();

        if (warp == 4) {
            // ---- MMA issuer: the whole warp runs the loop (uniform control flow, the operands stay in uniform registers), one
            //      elected lane issues; tile by tile, up to TC_RING tiles ahead of the DP warps.  Descriptors: only the low word
            //      (start address >> 4) moves.
            const uint32_t idesc = tcg::idesc_tf32(TC_LANES, TC_TILE);
            const uint32_t desc_hi = (uint32_t)((KCH * 128u) >> 4) | (1u << 14);
            const uint32_t b_lo0 = ((tcg::smem_u32(opB) & 0x3FFFFu) >> 4) | ((128u >> 4) << 16);
            const uint32_t a_lo0 = ((tcg::smem_u32(opA) & 0x3FFFFu) >> 4) | ((128u >> 4) << 16);
            for (int s = 0; s < R.max_rows; ++s) {
                const uint32_t a_lo = a_lo0 + a_buf * (uint32_t)((A_FLOATS * 4) >> 4);
                tcg::mbar_wait(bar_a + 8 * a_buf, a_par);
                uint32_t b_lo = b_lo0;
                for (int q = 0; q < n_tiles; ++q) {
                    tcg::mbar_wait(bar_ee + 8 * e_slot, e_par ^ 1);
                    tcg::fence_after();
                    if (tcg::elect_one()) {
#pragma unroll
                        for (int ks = 0; ks < K / 8; ++ks)
                            tcg::mma_tf32(tmem + e_slot * TC_TILE, ((uint64_t)desc_hi << 32) | (a_lo + ks * 16u), ((uint64_t)desc_hi << 32) | (b_lo + ks * 16u), idesc,
                                          ks > 0 ? 1u : 0u);
                        tcg::mma_commit(bar_ef + 8 * e_slot);
                    }
                    __syncwarp();
                    b_lo += (uint32_t)((TC_TILE / 8) * KCH * 128) >> 4;
                    if (++e_slot == TC_RING) { e_slot = 0; e_par ^= 1; }
                }
                a_par ^= a_buf; a_buf ^= 1;
            }
        } else {
            // ---- DP warps.  Registers are double buffered by half tiles: while the sixteen cells of one half are computed, the
            // exponents and the state of the next half are on their way from tensor memory.
            float acc = 0.f;                                   // H[i][last column of the strip]
            int istar = 0, isig = 0;                           // 1-based rows: last growth of H[i][m]; last growth above the float64 resolution
            float *bnd_w = g.bnd + R.bnd_base + tid;           // [row][128]
            // codes of the strips before this one: every strip but the last has ceil(strip_w / 32) tiles per row
            uint4 *tbp = g.tb + P.tb_base + (long long)strip * n_l * ((R.strip_w + TC_TILE - 1) / TC_TILE);
#ifdef TC_PROFILE
            long long t_a = 0, t_b = 0, t_wait = 0, t_row = 0, t_all0 = 0, t_all1 = 0;
#endif
            TC_T(t_all0);
            float a_next = (strip > 0) ? bnd_w[0] : 0.f;      // value entering the strip from the left, fetched a row ahead
            for (int s = 0; s < R.max_rows; ++s) {
                TC_T(t_a);
                if (s + 1 < R.max_rows) { stage_row(a_buf ^ 1); fetch_row(s + 2); }
                float a = a_next;
                if (strip > 0 && s + 1 < R.max_rows) a_next = bnd_w[(long long)(s + 1) * TC_LANES];
                const float th = acc * kappa;
                const bool live = have && s < n_l;
                uint32_t st_addr = lane_taddr + TC_STATE_COL;
                uint4 *tb_row = tbp + (long long)s * n_tiles;
                uint32_t eA[16], uA[16], eB[16], uB[16], w[3];
                unsigned word = 0;
                // first half of tile 0
                tcg::tmem_ld16(st_addr, uA);
#ifdef TC_PROFILE
                TC_T(t_b); t_row += t_b - t_a;
#endif
                tcg::mbar_wait(bar_ef + 8 * e_slot, e_par);
#ifdef TC_PROFILE
                TC_T(t_a); t_wait += t_a - t_b;
#endif
                tcg::fence_after();
                tcg::tmem_ld16(lane_taddr + e_slot * TC_TILE, eA);
                if (s == 0 && strip == 0) {                    // S[0][0] == 0: a zero region exists (k_trace emulates the stop state)
                    tcg::tmem_wait_ld();
                    if (have) g.pair_zflag[P.slot] = ex2_approx(__uint_as_float(eA[0])) == 0.f ? 1 : 0;
                }
                for (int q = 0; q < n_full; ++q) {
                    tcg::tmem_wait_ld();                                             // first half in registers
                    tcg::tmem_wait_st();                                             // uB has left for tensor memory (previous tile)
                    tcg::tmem_ld16(st_addr + 16, uB);
                    tcg::tmem_ld16(lane_taddr + e_slot * TC_TILE + 16, eB);
                    tc_cells16<0>(eA, uA, a, th, neg_eps, word, w);
                    tcg::tmem_st16(st_addr, uA);
                    tcg::tmem_wait_ld();                                             // second half in registers: the tile can be refilled
                    tcg::fence_before();
                    if (lane == 0) tcg::mbar_arrive(bar_ee + 8 * e_slot);
                    if (++e_slot == TC_RING) { e_slot = 0; e_par ^= 1; }
                    if (q + 1 < n_tiles) {
                        tcg::tmem_wait_st();                                         // uA has left for tensor memory
                        tcg::tmem_ld16(st_addr + 32, uA);
                        TC_T(t_b);
                        tcg::mbar_wait(bar_ef + 8 * e_slot, e_par);
        #ifdef TC_PROFILE
                TC_T(t_a); t_wait += t_a - t_b;
#endif
                        tcg::fence_after();
                        tcg::tmem_ld16(lane_taddr + e_slot * TC_TILE, eA);
                    }
                    tc_cells16<1>(eB, uB, a, th, neg_eps, word, w);
                    tcg::tmem_st16(st_addr + 16, uB);
                    if (live) tb_row[q] = make_uint4(w[0], w[1], w[2], 0u);
                    st_addr += TC_TILE;
                }
                if (half_last) {
                    tcg::tmem_wait_ld();
                    tcg::fence_before();
                    if (lane == 0) tcg::mbar_arrive(bar_ee + 8 * e_slot);
                    if (++e_slot == TC_RING) { e_slot = 0; e_par ^= 1; }
                    tc_cells16<0>(eA, uA, a, th, neg_eps, word, w);
                    tcg::tmem_st16(st_addr, uA);
                    if (live) tb_row[n_full] = make_uint4(w[0], word << 16, 0u, 0u);
                }
                a_buf ^= 1;
                TC_T(t_a);
                tcg::tmem_wait_st();
                // a = H[s][end of strip] - H[s-1][end of strip]
                if (!last_strip) bnd_w[(long long)s * TC_LANES] = a;
                acc += a;
                if (last_strip && live) {
                    if (a > 0.f) istar = s + 1;
                    if (a > th) isig = s + 1;
                    if (s == n_l - 1) {
                        g.pair_score[P.slot] = (double)acc;
                        g.pair_istar[P.slot] = istar | (istar != isig ? ISTAR_TIE : 0);
                    }
                }
#ifdef TC_PROFILE
                TC_T(t_b); t_row += t_b - t_a;
#endif
            }
            TC_T(t_all1);
#ifdef TC_PROFILE
            if (blockIdx.x == 0 && tid == 0 && g.prof) { atomicAdd((unsigned long long *)g.prof, (unsigned long long)t_wait); atomicAdd((unsigned long long *)g.prof + 1, (unsigned long long)t_row); atomicAdd((unsigned long long *)g.prof + 2, (unsigned long long)(t_all1 - t_all0)); atomicAdd((unsigned long long *)g.prof + 3, (unsigned long long)R.max_rows * n_tiles); }
#endif
        }
    }
  }
    tcg::fence_before();
    __syncthreads();
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TC_TMEM_COLS));
}

// ------------------------------------------------------------------------------------------------------------
// k_trace_tc: k_trace (crt_kernels.cuh) on the code layout of k_fill1_tc.  One thread per pair (= per partner record of the
// batch): start cell = first row-major maximum (dynamic_time_warping.py:241-247), walk with priority diag > left > up
// (:255-277), then the shared tail (Kabsch, by-products).  Codes of pair p: [strip][row][tile] uint4, 3 bits per column in push
// order (S attains the maximum, left attains it, suspect), MSB first.  A pair with a zero region (S[0][0] == 0: the reference's
// walk can stop inside the matrix) is handed to the float64 re-run: the exponents of the tensor-core tile are not bit-identical
// to a scalar re-evaluation, so the stop state is not emulated here.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TRACE_THREADS, CRT_TRACE_MINB) k_trace_tc(TraceArgs a, int n_part)
{
    const int gid = blockIdx.x * TRACE_THREADS + threadIdx.x;
    if (gid >= n_part) return;
    const TcPartner P = a.tc_partners[gid];
    const TcRound R = a.tc_rounds[P.round];
    const int pair = P.slot, n = P.n, m = R.m;
    const double *A = a.coords + (long long)P.row_base * 3;
    const double *B = a.coords + (long long)R.col_base * 3;
    const double *ceni = a.centroid + (long long)P.row_chain * 3, *cenj = a.centroid + (long long)R.col_chain * 3;
    short2 *path = a.path + P.path_base;
    const int tiles_full = (R.strip_w + TC_TILE - 1) / TC_TILE;
    const uint4 *tbp = a.tb + P.tb_base;

    // position of the walk: strip, column inside the strip, 0-based row; nt = tiles per row of the current strip
    int w_strip = 0, w_c = 0, w_row = 0, nt = tiles_full;
    auto strip_tiles = [&](int strip) {
        const int w_cols = min(R.strip_w, ((m - strip * R.strip_w + 15) / 16) * 16);
        return (w_cols + TC_TILE - 1) / TC_TILE;
    };
    auto seek = [&](int i, int j) {
        w_strip = (j - 1) / R.strip_w;
        w_c = (j - 1) - w_strip * R.strip_w;
        w_row = i - 1;
        nt = strip_tiles(w_strip);
    };
    auto col_left = [&]() {
        if (--w_c < 0) { --w_strip; w_c = R.strip_w - 1; nt = tiles_full; }
    };
    long long cidx = -1;
    uint4 cw = make_uint4(0, 0, 0, 0);
    unsigned tie = 0;
    auto code = [&]() -> unsigned {                       // (H != diag + S) << 1 | (H != left) of the current cell
        const long long idx = ((long long)w_strip * n * tiles_full) + (long long)w_row * nt + (w_c >> 5);
        if (idx != cidx) {
            cw = tbp[idx]; cidx = idx;
            if (w_row >= 2) {                             // the walk moves up and to the left
                prefetch_tb(tbp + idx - nt);
                prefetch_tb(tbp + idx - 2 * nt);
                if ((w_c & 31) < 2 && (w_c >> 5) > 0) prefetch_tb(tbp + idx - nt - 1);
            }
        }
        const int p = 3 * (w_c & 31);
        const unsigned hi = p < 32 ? cw.x : (p < 64 ? cw.y : cw.z), lo = p < 32 ? cw.y : (p < 64 ? cw.z : 0u);
        const unsigned raw = __funnelshift_l(lo, hi, p & 31) >> 29;      // (S attains) << 2 | (left attains) << 1 | suspect
        // suspect cell, or S and the left value attain the maximum together (an exact fp32 tie the reference may not share)
        tie |= (raw & 1u) | (raw >= 6u ? 1u : 0u);
        return ((raw ^ 6u) >> 1) & 3u;
    };

    const bool zreg = a.pair_zflag[pair] != 0;
    int wi = 0, wj = 0;
    auto is_zero_cell = [&](int i, int j) -> bool {
        if (wi > 0 && wi <= i && wj <= j) return false;
        for (int ii = 1; ii <= i; ++ii)
            for (int jj = 1; jj <= j; ++jj)
                if (!s1_is_zero(a, (long long)P.row_base + ii - 1, (long long)R.col_base + jj - 1)) { wi = ii; wj = jj; return false; }
        return true;
    };

    int i = a.pair_istar[pair], j = m;
    int st = 0, len = 0, c = 0;
    if (i & ISTAR_TIE) { tie = 1; i &= ~ISTAR_TIE; }
    if (zreg) tie = 1;
    if (i <= 0) {
        st |= 2;                      // CRT_ST_NO_POSITIVE
    } else {
        seek(i, j);
        while (j > 1 && (code() & 1u) == 0u) { --j; col_left(); }      // first column of row i* that attains the maximum
        while (i > 0 && j > 0) {
            if (zreg && is_zero_cell(i, j)) break;
            const unsigned cd = code();
            const bool diag = (cd & 2u) == 0u, left = !diag && (cd & 1u) == 0u;
            const bool di = diag || !left, dj = diag || left;
            i -= di ? 1 : 0; j -= dj ? 1 : 0; w_row -= di ? 1 : 0;
            if (dj) col_left();
            path[len++] = make_short2(di ? (short)i : (short)-1, dj ? (short)j : (short)-1);
            c += diag ? 1 : 0;
        }
    }
    trace_tail(a, pair, path, len, c, st, tie, n, m, A, B, ceni, cenj);
}

}  // namespace crt
