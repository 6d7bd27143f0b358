// crt_kernels.cuh -- device side of the caretta pair path for sm_100a.
//
// Data-parallel structure (DESIGN.md has the long form):
//   * a UNIT is one column chain j (the reference's seq2, multiple_alignment.py:164-169) and a run of consecutive
//     row chains i0..i1 (seq1).  One warp owns a unit.  Lane l keeps C consecutive columns of chain j in
//     REGISTERS (their feature vectors and the previous DP row), rows stream through the 32 lanes as a systolic
//     array: at step t lane l works on row t-l, the value crossing the strip boundary moves to lane l+1 with one
//     warp shuffle.  Row chains are streamed back to back, so the anti-diagonal wavefront never drains inside a unit.
//   * the Gaussian score S[a,b] (score_functions.py:6-11) is computed on the fly from the row record and the
//     lane's column registers; no n x m matrix ever exists in memory.
//   * stage 1 (Smith-Waterman on shape tensors + traceback, dynamic_time_warping.py:225-278) writes 2 bits per
//     cell -- (h != diag+S, h != left) -- as one coalesced 128-bit store per lane per 4 rows.
//   * k_trace walks those bits per pair (first row-major maximum, diag > left > up priority), accumulates the
//     Kabsch sums on the matched residues (superposition_functions.py:6-35), does the 3x3 Jacobi SVD with the
//     reflection fix, and writes the transformed row coordinates for stage 2.
//   * stage 2 (smith_waterman_score on superposed CA coordinates, dynamic_time_warping.py:204-222) is the same
//     systolic fill without traceback.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace crt {

constexpr unsigned FULL = 0xffffffffu;
constexpr int WARPS_PER_CTA = 4;
constexpr int ISTAR_TIE = 1 << 30;   // pair_istar flag: rows after the start row grew by less than the float64 resolution of H
constexpr int ST_TIE = 8;            // = CRT_ST_TIE: the fp32 walk met a decision the reference's float64 DP may take differently

struct Unit {
    long long row_base;    // packed residue index of the first row of the stream
    long long tb_base;     // traceback buffer, uint4 index
    long long rows2_base;  // stage-2 row records (batch-local index)
    long long bnd_base;    // strip boundary buffer (elements), multi-strip units only
    long long path_base;   // path buffer, short2 index
    int G;                 // rows in the stream (sum of the row chains' lengths)
    int m;                 // columns (length of chain j)
    int col_base;          // packed residue index of chain j
    int col_chain;         // j
    int row_chain0;        // i0
    int n_pairs;           // number of row chains
    int pair_base;         // index of pair (i0, j) in the run's result arrays
    int n_strips;          // ceil(m / (32*C))
    int path_stride;       // path entries reserved per pair
    int tchunks;           // 128-bit traceback chunks per lane per strip = ceil((G + 31) / 4)
    int dense_base;        // pairs of the batch's earlier units (k_trace maps threads densely onto pairs)
    long long s_base;      // node contexts: first element of the unit's precomputed score matrix [G][m] (crt_node_fill.cuh)
};

// per-residue row meta: bit0 = first residue of its chain, bit1 = last, bits 2.. = chain index
__host__ __device__ inline int make_meta(int chain, bool first, bool last) { return (chain << 2) | (last ? 2 : 0) | (first ? 1 : 0); }

__device__ __forceinline__ float ex2_approx(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ------------------------------------------------------------------------------------------------------------
// Score policies.  Row = what a lane loads every step, Col = what it keeps in registers per owned column.
// ------------------------------------------------------------------------------------------------------------

// stage 1, fp64 parity: the reference's arithmetic, score_functions.py:11 -- strict left-to-right sum of
// (a-b)*(a-b) with separate roundings (no FMA), then exp((-gamma) * acc).  Records: D doubles, zero padded
// beyond d (adds exact zeros), meta in a separate int array.
template <int D_>
struct P1F64 {
    typedef double T;
    static constexpr bool ROWS2 = false;
    static constexpr int D = D_;
    struct Row { double a[D]; int meta; };
    struct Col { double b[D]; };
    struct Args { const double *rec; const int *meta; double neg_gamma; };
    __device__ static __forceinline__ Row load_row(const Args &g, long long idx)
    {
        Row r;
        const double2 *p = reinterpret_cast<const double2 *>(g.rec + idx * D);
#pragma unroll
        for (int q = 0; q < D / 2; ++q) { double2 v = __ldg(p + q); r.a[2 * q] = v.x; r.a[2 * q + 1] = v.y; }
        r.meta = __ldg(g.meta + idx);
        return r;
    }
    __device__ static __forceinline__ Col load_col(const Args &g, long long idx)
    {
        Col c;
#pragma unroll
        for (int k = 0; k < D; ++k) c.b[k] = g.rec[idx * D + k];
        return c;
    }
    __device__ static __forceinline__ Col pad_col()
    {
        Col c;
#pragma unroll
        for (int k = 0; k < D; ++k) c.b[k] = 1e300;      // (a - 1e300)^2 = inf, exp(-inf) = 0
        return c;
    }
    __device__ static __forceinline__ double score(const Args &g, const Row &r, const Col &c)
    {
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            double t = __dsub_rn(r.a[k], c.b[k]);
            acc = __dadd_rn(acc, __dmul_rn(t, t));
        }
        return exp(__dmul_rn(g.neg_gamma, acc));
    }
};

// stage 2, fp64 parity: rows = (a - mean(common_1)) R^T + mean(common_2), columns = raw coordinates.
struct P2F64 {
    typedef double T;
    static constexpr bool ROWS2 = true;
    struct Row { double x, y, z; int meta; };
    struct Col { double x, y, z; };
    struct Args { const double *rows; const double *cols; double neg_gamma; };
    __device__ static __forceinline__ Row load_row(const Args &g, long long idx)
    {
        const double2 *p = reinterpret_cast<const double2 *>(g.rows + idx * 4);
        double2 v0 = __ldg(p), v1 = __ldg(p + 1);
        Row r; r.x = v0.x; r.y = v0.y; r.z = v1.x; r.meta = (int)__double_as_longlong(v1.y);
        return r;
    }
    __device__ static __forceinline__ Col load_col(const Args &g, long long idx)
    {
        Col c; c.x = g.cols[idx * 3]; c.y = g.cols[idx * 3 + 1]; c.z = g.cols[idx * 3 + 2];
        return c;
    }
    __device__ static __forceinline__ Col pad_col() { Col c; c.x = c.y = c.z = 1e300; return c; }
    __device__ static __forceinline__ double score(const Args &g, const Row &r, const Col &c)
    {
        double dx = __dsub_rn(r.x, c.x), dy = __dsub_rn(r.y, c.y), dz = __dsub_rn(r.z, c.z);
        double acc = __dmul_rn(dx, dx);
        acc = __dadd_rn(acc, __dmul_rn(dy, dy));
        acc = __dadd_rn(acc, __dmul_rn(dz, dz));
        return exp(__dmul_rn(g.neg_gamma, acc));
    }
};

template <typename T> __device__ __forceinline__ T max3(T a, T b, T c);
template <> __device__ __forceinline__ float max3<float>(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }
template <> __device__ __forceinline__ double max3<double>(double a, double b, double c)
{
    double t = a > b ? a : b;     // all operands are finite and >= 0 on this path
    return t > c ? t : c;
}
__device__ __forceinline__ float shfl_up1(float v) { return __shfl_up_sync(FULL, v, 1); }
__device__ __forceinline__ double shfl_up1(double v) { return __shfl_up_sync(FULL, v, 1); }

struct FillOut {
    uint4 *tb;           // traceback bits (stage 1)
    int *pair_istar;     // 1-based row of the first row-major maximum (0 = no positive cell)
    int *pair_zflag;     // 1 if S[0][0] == 0 (a zero region exists, k_trace handles the stop state)
    double *pair_score;  // stage 1: H[n][m] of the tensor SW; stage 2: the pair score
    void *bnd;           // strip boundary values
    const int *skip_status;   // stage 2 of the fp32 mode: pairs whose status carries ST_TIE are not written (the float64 re-run owns their
                              // slots and may already be running); nullptr = write every pair
};

// ------------------------------------------------------------------------------------------------------------
// Systolic Smith-Waterman fill, gap = 0.
//   DIFF = false: absolute form  h = max(H[i-1][j-1] + S, H[i][j-1], H[i-1][j])       (the reference's expression)
//   DIFF = true : difference form d = max(S, a, b), u' = d - a, v' = d - b with a = H[i][j-1]-H[i-1][j-1] (vertical
//                 difference of the left neighbour) and b = H[i-1][j]-H[i-1][j-1] (horizontal difference of the upper
//                 neighbour).  Every quantity stays in [0, 1], so fp32 keeps ~100x more absolute resolution than
//                 the absolute form; the decisions d==S / d==a are the same decisions in exact arithmetic.
//   CODES: emit the 2-bit traceback codes and the start row (stage 1).
// ------------------------------------------------------------------------------------------------------------
template <typename P, int C, bool DIFF, bool CODES, bool MULTI>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
k_fill(const Unit *__restrict__ units, int n_units, typename P::Args args, FillOut out)
{
    typedef typename P::T T;
    const int lane = threadIdx.x & 31;
    const int uidx = blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5);
    if (uidx >= n_units) return;
    const Unit u = units[uidx];
    const int G = u.G;
    const int steps4 = u.tchunks * 4;           // >= G + 31
    const long long rbase = P::ROWS2 ? u.rows2_base : u.row_base;
    T *bnd = MULTI ? reinterpret_cast<T *>(out.bnd) + u.bnd_base : nullptr;

    for (int strip = 0; strip < (MULTI ? u.n_strips : 1); ++strip) {
        typename P::Col col[C];
        const int c0 = (strip * 32 + lane) * C;
#pragma unroll
        for (int c = 0; c < C; ++c)
            col[c] = (c0 + c < u.m) ? P::load_col(args, (long long)u.col_base + c0 + c) : P::pad_col();
        const bool last_strip = !MULTI || strip == u.n_strips - 1;

        T prev[C];                 // DIFF: horizontal differences u[i-1][j]; else H[i-1][j]
#pragma unroll
        for (int c = 0; c < C; ++c) prev[c] = T(0);
        T carry = T(0);            // value handed to lane+1: DIFF: v[i][cend]; else H[i][cend]
        T dsave = T(0);            // abs form: H[i-1][c0-1]
        T acc = T(0);              // DIFF: running H[i][m] on the last lane
        int istar = 0, r = 0;
        uint4 *tbp = CODES ? out.tb + u.tb_base + (long long)strip * u.tchunks * 32 + lane : nullptr;

        for (int t0 = 0; t0 < steps4; t0 += 4) {
            unsigned w[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int g = t0 + q - lane;
                const bool valid = (unsigned)g < (unsigned)G;
                const int gc = min(max(g, 0), G - 1);
                const typename P::Row row = P::load_row(args, rbase + gc);
                T in = shfl_up1(carry);
                if (lane == 0) {
                    in = T(0);
                    if (MULTI && strip > 0) in = bnd[gc];
                }
                const bool first = valid && (row.meta & 1);
                if (first) {
#pragma unroll
                    for (int c = 0; c < C; ++c) prev[c] = T(0);
                    dsave = T(0); acc = T(0); istar = 0; r = 0;
                }
                unsigned word = 0;
                T left = in;           // DIFF: a
                T diag = dsave;
                T s0 = T(0);
                bool grew = false;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const T s = P::score(args, row, col[c]);
                    if (c == 0) s0 = s;
                    if (DIFF) {
                        const T b = prev[c];
                        const T d = max3<T>(s, left, b);
                        if (CODES) word |= ((d != s ? 2u : 0u) | (d != left ? 1u : 0u)) << (2 * (C - 1 - c));
                        prev[c] = d - left;
                        left = d - b;
                    } else {
                        const T up = prev[c];
                        const T dg = diag + s;
                        const T h = max3<T>(dg, left, up);
                        if (CODES) word |= ((h != dg ? 2u : 0u) | (h != left ? 1u : 0u)) << (2 * (C - 1 - c));
                        if (c == C - 1) grew = h > up;
                        diag = up;
                        prev[c] = h;
                        left = h;
                    }
                }
                carry = left;
                dsave = in;
                ++r;
                if (DIFF) { grew = left > T(0); acc += left; }
                if (CODES && grew) istar = r;          // last row whose H[i][m] exceeds H[i-1][m] (meaningful on lane 31)
                w[q] = word;
                if (valid) {
                    if (MULTI && !last_strip && lane == 31) bnd[g] = carry;
                    if (CODES && lane == 0 && strip == 0 && (row.meta & 1))
                        out.pair_zflag[u.pair_base + (row.meta >> 2) - u.row_chain0] = (s0 == T(0)) ? 1 : 0;
                    if (last_strip && lane == 31 && (row.meta & 2)) {
                        const int pidx = u.pair_base + (row.meta >> 2) - u.row_chain0;
                        out.pair_score[pidx] = DIFF ? (double)acc : (double)carry;
                        if (CODES) out.pair_istar[pidx] = istar;
                    }
                }
            }
            if (CODES) tbp[(long long)(t0 >> 2) * 32] = make_uint4(w[0], w[1], w[2], w[3]);
        }
        if (MULTI) __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------------------
// 3x3 SVD by one-sided Jacobi in fp64 and the Kabsch rotation with the reference's reflection fix
// (superposition_functions.py:27-33): R = U diag(1,1,sign) Vt for C = X2c^T X1c, row-vector convention x2 R ~ x1.
// ------------------------------------------------------------------------------------------------------------
__device__ inline void kabsch_rotation(const double Cm[9], double R[9])
{
    double W[9], V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
#pragma unroll
    for (int q = 0; q < 9; ++q) W[q] = Cm[q];
    for (int sweep = 0; sweep < 40; ++sweep) {
        double off = 0.0;
#pragma unroll
        for (int pq = 0; pq < 3; ++pq) {
            const int p = (pq == 2) ? 1 : 0, q = (pq == 0) ? 1 : 2;
            double al = 0, be = 0, ga = 0;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                al += W[r * 3 + p] * W[r * 3 + p];
                be += W[r * 3 + q] * W[r * 3 + q];
                ga += W[r * 3 + p] * W[r * 3 + q];
            }
            const double lim = sqrt(al * be);
            if (ga == 0.0 || fabs(ga) <= 1e-300 || fabs(ga) <= 2.2e-16 * lim) continue;
            off = fmax(off, fabs(ga) / (lim > 0 ? lim : 1.0));
            const double zeta = (be - al) / (2.0 * ga);
            const double tt = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
            const double cs = 1.0 / sqrt(1.0 + tt * tt), sn = cs * tt;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const double wp = W[r * 3 + p], wq = W[r * 3 + q];
                W[r * 3 + p] = cs * wp - sn * wq; W[r * 3 + q] = sn * wp + cs * wq;
                const double vp = V[r * 3 + p], vq = V[r * 3 + q];
                V[r * 3 + p] = cs * vp - sn * vq; V[r * 3 + q] = sn * vp + cs * vq;
            }
        }
        if (off == 0.0) break;
    }
    double nrm[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) nrm[q] = sqrt(W[q] * W[q] + W[3 + q] * W[3 + q] + W[6 + q] * W[6 + q]);
    // order columns by decreasing singular value (o0, o1, o2)
    int o0 = 0, o1 = 1, o2 = 2;
    if (nrm[o1] > nrm[o0]) { int t = o0; o0 = o1; o1 = t; }
    if (nrm[o2] > nrm[o0]) { int t = o0; o0 = o2; o2 = t; }
    if (nrm[o2] > nrm[o1]) { int t = o1; o1 = o2; o2 = t; }
    double U[3][3], Vc[3][3];      // U[q] = q-th left singular vector, Vc[q] = q-th right singular vector
    const int ord[3] = {o0, o1, o2};
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        const int o = ord[q];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            Vc[q][r] = V[r * 3 + o];
            U[q][r] = nrm[o] > 0 ? W[r * 3 + o] / nrm[o] : 0.0;
        }
    }
    const double s0 = nrm[o0], s1 = nrm[o1], s2 = nrm[o2];
    const double tiny = 1e-14 * (s0 > 0 ? s0 : 1.0);
    if (!(s2 > tiny)) {
        if (!(s1 > tiny)) {
            if (!(s0 > 0)) { U[0][0] = 1; U[0][1] = 0; U[0][2] = 0; }
            int mn = 0;
            if (fabs(U[0][1]) < fabs(U[0][mn])) mn = 1;
            if (fabs(U[0][2]) < fabs(U[0][mn])) mn = 2;
            double ax[3] = {0, 0, 0};
            ax[mn] = 1.0;
            const double dp = U[0][mn];
            double nn = 0;
#pragma unroll
            for (int r = 0; r < 3; ++r) { U[1][r] = ax[r] - dp * U[0][r]; nn += U[1][r] * U[1][r]; }
            nn = sqrt(nn);
#pragma unroll
            for (int r = 0; r < 3; ++r) U[1][r] /= nn;
        }
        U[2][0] = U[0][1] * U[1][2] - U[0][2] * U[1][1];
        U[2][1] = U[0][2] * U[1][0] - U[0][0] * U[1][2];
        U[2][2] = U[0][0] * U[1][1] - U[0][1] * U[1][0];
    }
    // det of the matrices whose COLUMNS are U[q] (resp. rows of Vt are Vc[q]); det is transpose invariant
    const double detU = U[0][0] * (U[1][1] * U[2][2] - U[1][2] * U[2][1]) - U[0][1] * (U[1][0] * U[2][2] - U[1][2] * U[2][0])
                      + U[0][2] * (U[1][0] * U[2][1] - U[1][1] * U[2][0]);
    const double detV = Vc[0][0] * (Vc[1][1] * Vc[2][2] - Vc[1][2] * Vc[2][1]) - Vc[0][1] * (Vc[1][0] * Vc[2][2] - Vc[1][2] * Vc[2][0])
                      + Vc[0][2] * (Vc[1][0] * Vc[2][1] - Vc[1][1] * Vc[2][0]);
    const double sg = (detU * detV < 0) ? -1.0 : 1.0;
    // R[a][b] = sum_q Umat[a][q] * Vt[q][b] = sum_q U[q][a] * Vc[q][b], last term sign-flipped on reflection
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b)
            R[a * 3 + b] = U[0][a] * Vc[0][b] + U[1][a] * Vc[1][b] + sg * U[2][a] * Vc[2][b];
}

// ------------------------------------------------------------------------------------------------------------
// k_trace: one thread per pair.  Start cell = first row-major maximum (dynamic_time_warping.py:241-247), walk with
// priority diag > left > up (:255-277), common positions (helper.py:12-42), Kabsch (superposition_functions.py:6-60),
// by-products RMSD / TM (score_functions.py:14-19, multiple_alignment.py:59-70), stage-2 row records.
// ------------------------------------------------------------------------------------------------------------
struct TcRound;
struct TcPartner;
// fp32 production mode, where the walk of a pair FIRST met a decision the reference may take differently (windowed re-run):
// everything the walk did before that cell is the reference's, so the float64 re-run only fills rows 1..i x columns 1..j (H there
// does not depend on the rest), resumes the walk AT cell (i, j) and keeps the first `len` path entries (saved at `off` of the pool).
// i = 0: no window (the start row itself was in doubt, a zero-region pair, or the pool was full): the whole pair is re-run.
struct TieInfo { int i, j, len, c; long long off; };
struct TraceArgs {
    const Unit *units;
    const TcRound *tc_rounds;     // k_trace_tc: the rounds / partners of the batch (crt_fill_tc.cuh)
    const TcPartner *tc_partners;
    const uint4 *tb;
    const int *pair_istar;
    const int *pair_zflag;
    const long long *offsets;     // [N+1]
    const double *coords;         // raw CA coordinates [sumL,3]
    const double *centroid;       // per-chain centroid [N,3] (fp32 records are relative to it)
    short2 *path;
    int *path_len;                // per pair
    double *rmsd, *tm;
    int *ncommon, *status;
    double *xform;                // [pairs, XF]: R (9), mean1 (3), mean2 (3), superpose flag
    const int *meta;              // per-residue row meta (chain index, first/last flags)
    void *rows2;                  // float4[] (fp32) or double[4][] (fp64)
    // for the exact zero-region test (stop state of the reference traceback)
    const float *rec32; int rs32; int d32;
    const double *rec64; int d64; double neg_gamma_t;
    float scale2;                 // sqrt(gamma_c * log2 e) (fp32 only)
    int precision;                // 0 fp64, 1 fp32: the stage-1 fill that ran (records of the exact zero test)
    int rows2_f32;                // format of the stage-2 row records: 1 = float4 (fp32 stage 2), 0 = double[4]
    int skip_byproducts;          // 1: no RMSD / TM pass over the path (node contexts only use the transform)
    int status_or;                // bits OR-ed into every status written (the overlapped float64 re-run keeps ST_TIE up until the run ends)
    // windowed re-run: the main fp32 traceback writes tie_out / the pool, the float64 traceback of the re-run reads tie_in
    TieInfo *tie_out; const TieInfo *tie_in;
    short2 *tie_pool; unsigned long long *tie_pool_used; unsigned long long tie_pool_cap;
};

__device__ inline bool s1_is_zero(const TraceArgs &a, long long ri, long long ci)
{
    if (a.precision == 1) {
        // same operation order as the stage-1 fill (rbf_row_v2): e = A_col; e = fma(r_k, c_k, e) for k = 0..D-1; e += A_row
        const float *x = a.rec32 + ri * a.rs32, *y = a.rec32 + ci * a.rs32;
        float e = y[a.d32];
        for (int k = 0; k < a.d32; ++k) e = __fmaf_rn(x[k], y[k], e);
        return ex2_approx(__fadd_rn(e, x[a.d32])) == 0.f;
    }
    const double *x = a.rec64 + ri * a.d64, *y = a.rec64 + ci * a.d64;
    double acc = 0.0;
    for (int k = 0; k < a.d64; ++k) { double t = __dsub_rn(x[k], y[k]); acc = __dadd_rn(acc, __dmul_rn(t, t)); }
    return exp(__dmul_rn(a.neg_gamma_t, acc)) == 0.0;
}

constexpr int XF = 16;                // doubles per pair in the transform array: R[9], m1[3], m2[3], superpose flag

// Passes 2 and 3 of the traceback kernels: Kabsch on the matched residues of a walked path (superposition_functions.py:6-60),
// transform record for k_rows2, by-products RMSD / TM (score_functions.py:14-19, multiple_alignment.py:59-70).
// path: len entries as walked (descending), (-1, j) / (i, -1) = gaps; c = matched residues; st = status bits so far.
__device__ __forceinline__ void trace_tail(const TraceArgs &a, int pair, const short2 *path, int len, int c, int st, unsigned tie, int n, int m,
                                           const double *A, const double *B, const double *ceni, const double *cenj)
{
    const double ca0 = ceni[0], ca1 = ceni[1], ca2 = ceni[2], cb0 = cenj[0], cb1 = cenj[1], cb2 = cenj[2];
    a.path_len[pair] = len;
    a.ncommon[pair] = c;
    if (tie) st |= ST_TIE;

    // ---- pass 2: moments of the matched residues, relative to the chains' own centroids (translation does not change
    // the covariance; it keeps the raw-moment form well conditioned).  Same summation order as the walk (descending).
    double s1x = 0, s1y = 0, s1z = 0, s2x = 0, s2y = 0, s2z = 0;
    double cxx = 0, cxy = 0, cxz = 0, cyx = 0, cyy = 0, cyz = 0, czx = 0, czy = 0, czz = 0;
    const bool superpose = c > 3;
    if (superpose) {
#pragma unroll 4
        for (int q = 0; q < len; ++q) {
            const short2 e = path[q];
            if (e.x < 0 || e.y < 0) continue;
            const double x1 = A[e.x * 3] - ca0, y1 = A[e.x * 3 + 1] - ca1, z1 = A[e.x * 3 + 2] - ca2;
            const double x2 = B[e.y * 3] - cb0, y2 = B[e.y * 3 + 1] - cb1, z2 = B[e.y * 3 + 2] - cb2;
            s1x += x1; s1y += y1; s1z += z1; s2x += x2; s2y += y2; s2z += z2;
            cxx += x2 * x1; cxy += x2 * y1; cxz += x2 * z1;
            cyx += y2 * x1; cyy += y2 * y1; cyz += y2 * z1;
            czx += z2 * x1; czy += z2 * y1; czz += z2 * z1;
        }
    }

    double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    double m1[3] = {0, 0, 0}, m2[3] = {0, 0, 0};
    if (!superpose) st |= 1;          // CRT_ST_FEW_COMMON: multiple_alignment.py:337-342
    if (superpose) {
        const double inv = 1.0 / (double)c;
        const double p1[3] = {s1x * inv, s1y * inv, s1z * inv}, p2[3] = {s2x * inv, s2y * inv, s2z * inv};
        double Cm[9];                 // sum (x2 - mean2)(x1 - mean1)^T = raw moments - c * mean2 mean1^T
        Cm[0] = cxx - s2x * p1[0]; Cm[1] = cxy - s2x * p1[1]; Cm[2] = cxz - s2x * p1[2];
        Cm[3] = cyx - s2y * p1[0]; Cm[4] = cyy - s2y * p1[1]; Cm[5] = cyz - s2y * p1[2];
        Cm[6] = czx - s2z * p1[0]; Cm[7] = czy - s2z * p1[1]; Cm[8] = czz - s2z * p1[2];
        kabsch_rotation(Cm, R);
        m1[0] = p1[0] + ca0; m1[1] = p1[1] + ca1; m1[2] = p1[2] + ca2;
        m2[0] = p2[0] + cb0; m2[1] = p2[1] + cb1; m2[2] = p2[2] + cb2;
    }
    // translation of apply_rotran: t = m1 - m2 R
    double tr[3];
    for (int b = 0; b < 3; ++b) tr[b] = m1[b] - (m2[0] * R[b] + m2[1] * R[3 + b] + m2[2] * R[6 + b]);
    double *xf = a.xform + (long long)pair * XF;
    for (int q = 0; q < 9; ++q) xf[q] = R[q];
    for (int q = 0; q < 3; ++q) { xf[9 + q] = m1[q]; xf[12 + q] = m2[q]; }
    xf[15] = superpose ? 1.0 : 0.0;
    // ---- pass 3: by-products over the matched residues, ascending residue order like the reference's sums
    double rmsd = 0.0, tm = 0.0;
    if (c >= 1 && !a.skip_byproducts) {
        const double d1 = 1.24 * (double)(n - 15) / 3 - 1.8, d2 = 1.24 * (double)(m - 15) / 3 - 1.8;
        const double id1 = 1.0 / d1, id2 = 1.0 / d2;
        double ss = 0.0, t1 = 0.0, t2 = 0.0;
#pragma unroll 2
        for (int q = len - 1; q >= 0; --q) {
            const short2 e = path[q];
            if (e.x < 0 || e.y < 0) continue;
            double sm = 0.0;
            const double *y = B + e.y * 3;
            const double y0 = y[0], y1 = y[1], y2 = y[2];
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                const double yr = superpose ? (y0 * R[b] + y1 * R[3 + b] + y2 * R[6 + b]) + tr[b] : (b == 0 ? y0 : (b == 1 ? y1 : y2));
                const double df = A[e.x * 3 + b] - yr;
                ss += df * df;
                sm += df;
            }
            // 1 / (1 + (sm / d)^2) for both normalisations by ONE division (a float64 division is ~30 instructions, and this
            // pass ran four per matched residue): x_k = 1 + (sm / d_k)^2, r = 1 / (x_1 x_2), 1 / x_1 = r x_2.  A few ulp
            // from tm_score's own operation order (multiple_alignment.py:60-70); the by-products are tested to 1e-8.
            const double q1 = sm * id1, q2 = sm * id2;
            const double x1 = 1 + q1 * q1, x2 = 1 + q2 * q2;
            const double r = 1 / (x1 * x2);
            t1 += r * x2;
            t2 += r * x1;
        }
        rmsd = sqrt(ss / (double)c);
        t1 = (1.0 / (double)n) * t1;
        t2 = (1.0 / (double)m) * t2;
        tm = t1 > t2 ? t1 : t2;
    }
    a.rmsd[pair] = rmsd;
    a.tm[pair] = tm;
    a.status[pair] = st | a.status_or;
}

// Passes 2 and 3 of ONE pair by its warp (k_trace_w: small batches -- tree levels, the float64 re-run, small inputs --, where a
// thread per pair leaves the device empty and each pass is ~n + m dependent round trips of one thread).  Lane l takes the steps
// l, l + 32, ... of the path, two (pass 3: four) at a time and branch-free, so the loads of a warp fall into a few sectors and
// several are in flight; 15 + 3 shuffle-tree sums; lane 0 keeps the 3 x 3 SVD and writes the results.  The sums run in another
// order than trace_tail's (differences of a few ulp in the rotation: the goldens hold to 1e-9 / 1e-8, the alignments are unchanged).
__device__ __forceinline__ double warp_sum_t(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

__device__ __forceinline__ void trace_tail_warp(const TraceArgs &a, int pair, const short2 *path, int len, int c, int st, unsigned tie, int n, int m,
                                                const double *A, const double *B, const double *ceni, const double *cenj)
{
    const int lane = threadIdx.x & 31;
    const double ca0 = ceni[0], ca1 = ceni[1], ca2 = ceni[2], cb0 = cenj[0], cb1 = cenj[1], cb2 = cenj[2];
    if (tie) st |= ST_TIE;
    const bool superpose = c > 3;
    double mom[15];
#pragma unroll
    for (int k = 0; k < 15; ++k) mom[k] = 0.0;
    if (superpose) {
        for (int q0 = lane; q0 < len; q0 += 64) {
            double v[2][6], w[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int q = q0 + 32 * h;
                const short2 e = q < len ? path[q] : make_short2(-1, -1);
                const bool ok = e.x >= 0 && e.y >= 0;
                const int ia = ok ? e.x * 3 : 0, ib = ok ? e.y * 3 : 0;
                w[h] = ok ? 1.0 : 0.0;
                v[h][0] = A[ia]; v[h][1] = A[ia + 1]; v[h][2] = A[ia + 2];
                v[h][3] = B[ib]; v[h][4] = B[ib + 1]; v[h][5] = B[ib + 2];
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const double x1 = (v[h][0] - ca0) * w[h], y1 = (v[h][1] - ca1) * w[h], z1 = (v[h][2] - ca2) * w[h];
                const double x2 = (v[h][3] - cb0) * w[h], y2 = (v[h][4] - cb1) * w[h], z2 = (v[h][5] - cb2) * w[h];
                mom[0] += x1; mom[1] += y1; mom[2] += z1; mom[3] += x2; mom[4] += y2; mom[5] += z2;
                mom[6] += x2 * x1; mom[7] += x2 * y1; mom[8] += x2 * z1;
                mom[9] += y2 * x1; mom[10] += y2 * y1; mom[11] += y2 * z1;
                mom[12] += z2 * x1; mom[13] += z2 * y1; mom[14] += z2 * z1;
            }
        }
#pragma unroll
        for (int k = 0; k < 15; ++k) mom[k] = warp_sum_t(mom[k]);
    }
    double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    double m1[3] = {0, 0, 0}, m2[3] = {0, 0, 0};
    if (!superpose) st |= 1;          // CRT_ST_FEW_COMMON: multiple_alignment.py:337-342
    if (superpose && lane == 0) {
        const double inv = 1.0 / (double)c;
        const double p1[3] = {mom[0] * inv, mom[1] * inv, mom[2] * inv}, p2[3] = {mom[3] * inv, mom[4] * inv, mom[5] * inv};
        double Cm[9];
        Cm[0] = mom[6] - mom[3] * p1[0]; Cm[1] = mom[7] - mom[3] * p1[1]; Cm[2] = mom[8] - mom[3] * p1[2];
        Cm[3] = mom[9] - mom[4] * p1[0]; Cm[4] = mom[10] - mom[4] * p1[1]; Cm[5] = mom[11] - mom[4] * p1[2];
        Cm[6] = mom[12] - mom[5] * p1[0]; Cm[7] = mom[13] - mom[5] * p1[1]; Cm[8] = mom[14] - mom[5] * p1[2];
        kabsch_rotation(Cm, R);
        m1[0] = p1[0] + ca0; m1[1] = p1[1] + ca1; m1[2] = p1[2] + ca2;
        m2[0] = p2[0] + cb0; m2[1] = p2[1] + cb1; m2[2] = p2[2] + cb2;
    }
    double tr[3];
#pragma unroll
    for (int b = 0; b < 3; ++b) tr[b] = m1[b] - (m2[0] * R[b] + m2[1] * R[3 + b] + m2[2] * R[6 + b]);
    if (lane == 0) {
        a.path_len[pair] = len;
        a.ncommon[pair] = c;
        double *xf = a.xform + (long long)pair * XF;
#pragma unroll
        for (int q = 0; q < 9; ++q) xf[q] = R[q];
#pragma unroll
        for (int q = 0; q < 3; ++q) { xf[9 + q] = m1[q]; xf[12 + q] = m2[q]; }
        xf[15] = superpose ? 1.0 : 0.0;
    }
    double rmsd = 0.0, tm = 0.0;
    if (c >= 1 && !a.skip_byproducts) {
#pragma unroll
        for (int k = 0; k < 9; ++k) R[k] = __shfl_sync(FULL, R[k], 0);
#pragma unroll
        for (int k = 0; k < 3; ++k) tr[k] = __shfl_sync(FULL, tr[k], 0);
        const double d1 = 1.24 * (double)(n - 15) / 3 - 1.8, d2 = 1.24 * (double)(m - 15) / 3 - 1.8;
        const double id1 = 1.0 / d1, id2 = 1.0 / d2;
        double ss = 0.0, t1 = 0.0, t2 = 0.0;
        for (int q0 = lane; q0 < len; q0 += 128) {
            double v[4][6], w[4];
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                const int q = q0 + 32 * h;
                const short2 e = q < len ? path[q] : make_short2(-1, -1);
                const bool ok = e.x >= 0 && e.y >= 0;
                const int ia = ok ? e.x * 3 : 0, ib = ok ? e.y * 3 : 0;
                w[h] = ok ? 1.0 : 0.0;
                v[h][0] = A[ia]; v[h][1] = A[ia + 1]; v[h][2] = A[ia + 2];
                v[h][3] = B[ib]; v[h][4] = B[ib + 1]; v[h][5] = B[ib + 2];
            }
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                double sm = 0.0, sq = 0.0;
#pragma unroll
                for (int b = 0; b < 3; ++b) {
                    // (R = identity, t = 0 without a superposition: the products and sums are then exact)
                    const double yr = (v[h][3] * R[b] + v[h][4] * R[3 + b] + v[h][5] * R[6 + b]) + tr[b];
                    const double df = v[h][b] - yr;
                    sq += df * df;
                    sm += df;
                }
                const double q1 = sm * id1, q2 = sm * id2;            // one division for both terms, as in trace_tail
                const double x1 = 1 + q1 * q1, x2 = 1 + q2 * q2;
                const double r = w[h] / (x1 * x2);
                ss += w[h] * sq;
                t1 += r * x2;
                t2 += r * x1;
            }
        }
        ss = warp_sum_t(ss); t1 = warp_sum_t(t1); t2 = warp_sum_t(t2);
        rmsd = sqrt(ss / (double)c);
        t1 = (1.0 / (double)n) * t1;
        t2 = (1.0 / (double)m) * t2;
        tm = t1 > t2 ? t1 : t2;
    }
    if (lane == 0) {
        a.rmsd[pair] = rmsd;
        a.tm[pair] = tm;
        a.status[pair] = st | a.status_or;
    }
}

#ifndef CRT_TRACE_MINB
#define CRT_TRACE_MINB 6
#endif
constexpr int TRACE_THREADS = 128;    // threads (= pairs) per CTA of k_trace

__device__ __forceinline__ void prefetch_tb(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// One thread per pair, pairs mapped densely onto threads (Unit::dense_base is the batch-local prefix of n_pairs; the
// thread finds its unit by bisection).  Three passes per pair:
//   1. the walk: only the dependent chain  traceback word -> code -> next cell  and the path store; the words the walk
//      will need next (one and two chunk rows up, same lane and the lane to the left) are prefetched when a new chunk
//      is entered, so most dependent loads hit L1/L2 instead of HBM;
//   2. Kabsch moments over the matched residues, read back from the path (independent loads, unrolled);
//   3. the by-products (RMSD, TM) with the rotation applied.
// TIE3 = false: 2 bits per cell, (H != diag + S) << 1 | (H != left), written by the float64 k_fill.
// TIE3 = true : 3 bits per cell, (S attains the maximum) << 2 | (left attains it) << 1 | suspect, written by k_fill1_v4; a walk
//               that meets a suspect cell sets ST_TIE in the pair's status (the host re-runs those pairs in float64).
// WARP = true (k_trace_w below): a warp per pair; lane 0 walks, the warp does the passes over the path (trace_tail_warp).
template <int C, bool TIE3, bool WARP>
__device__ __forceinline__ void trace_pair(const TraceArgs &a, int n_units, int n_dense)
{
    const int gid = WARP ? (int)((blockIdx.x * TRACE_THREADS + threadIdx.x) >> 5) : (int)(blockIdx.x * TRACE_THREADS + threadIdx.x);
    if (gid >= n_dense) return;
    int lo = 0, hi = n_units - 1;             // last unit with dense_base <= gid
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (a.units[mid].dense_base <= gid) lo = mid; else hi = mid - 1;
    }
    const Unit u = a.units[lo];
    const int tid = gid - u.dense_base;
    const int ci = u.row_chain0 + tid;
    const int pair = u.pair_base + tid;
    const long long ro = a.offsets[ci];
    const int n = (int)(a.offsets[ci + 1] - ro), m = u.m;      // m: the unit's columns (a window of the re-run: fewer than the chain's)
    const int m_full = (int)(a.offsets[u.col_chain + 1] - a.offsets[u.col_chain]);
    const int g0 = (int)(ro - u.row_base);
    const double *A = a.coords + ro * 3;
    const double *B = a.coords + (long long)u.col_base * 3;
    const double *ceni = a.centroid + (long long)ci * 3, *cenj = a.centroid + (long long)u.col_chain * 3;
    short2 *path = a.path + u.path_base + (long long)tid * u.path_stride;
    constexpr int SW = 32 * C;

    // Position of the walk in the traceback layout, kept incrementally (no divisions in the loop): cell (i, j) 1-based
    // -> strip, lane l, column k inside the lane, unit row trow = g0 + i - 1; chunk = (trow + l) >> 2.
    const uint4 *tbu = a.tb + u.tb_base;
    int w_strip = 0, w_l = 0, w_k = 0, w_trow = 0;
    auto seek = [&](int i, int j) {
        const int c = j - 1;
        w_strip = c / SW;
        const int cc = c - w_strip * SW;
        w_l = cc / C; w_k = cc - w_l * C;
        w_trow = g0 + i - 1;
    };
    auto col_left = [&]() {
        if (--w_k < 0) { w_k = C - 1; if (--w_l < 0) { w_l = 31; --w_strip; } }
    };
    int cidx = -1;
    uint4 cw = make_uint4(0, 0, 0, 0);
    unsigned tie = 0;
    auto code = [&]() -> unsigned {                       // (nA << 1) | nB of the current cell
        const int t = w_trow + w_l;
        const int idx = (w_strip * u.tchunks + (t >> 2)) * 32 + w_l;
        if (idx != cidx) {
            cw = tbu[idx]; cidx = idx;
#ifndef CRT_TRACE_PF
#define CRT_TRACE_PF 3
#endif
            if (CRT_TRACE_PF > 0 && (t >> 2) >= 2) {      // the walk moves up and to the left
                prefetch_tb(tbu + idx - 64);
                if (CRT_TRACE_PF > 1 && w_l > 0) { prefetch_tb(tbu + idx - 33); if (CRT_TRACE_PF > 2) prefetch_tb(tbu + idx - 65); }
            }
        }
        const int q = t & 3;
        const unsigned w = q == 0 ? cw.x : (q == 1 ? cw.y : (q == 2 ? cw.z : cw.w));
        if (TIE3) {
            const unsigned raw = (w >> (3 * (C - 1 - w_k))) & 7u;
            // suspect cell, or S and the left value attain the maximum together (an exact fp32 tie the reference may not share)
            tie |= (raw & 1u) | (raw >= 6u ? 1u : 0u);
            return ((raw ^ 6u) >> 1) & 3u;
        }
        return (w >> (2 * (C - 1 - w_k))) & 3u;
    };

    // exact stop-state emulation: H[i][j] == 0 iff every S in [1..i] x [1..j] is 0; only possible when S[1][1] == 0
    const bool zreg = a.pair_zflag[pair] != 0;
    int wi = 0, wj = 0;               // witness: a nonzero S at (wi, wj) (1-based), 0 = none known
    auto is_zero_cell = [&](int i, int j) -> bool {
        if (wi > 0 && wi <= i && wj <= j) return false;
        for (int ii = 1; ii <= i; ++ii)
            for (int jj = 1; jj <= j; ++jj)
                if (!s1_is_zero(a, ro + ii - 1, (long long)u.col_base + jj - 1)) { wi = ii; wj = jj; return false; }
        return true;
    };

    // ---- pass 1: the walk
    int i = a.pair_istar[pair], j = m;
    int st = 0, len = 0, c = 0;
    // windowed re-run (float64 traceback of a pair the fp32 walk marked): resume at the marked cell behind the saved prefix
    TieInfo ti{0, 0, 0, 0, 0};
    if (!TIE3 && a.tie_in) ti = a.tie_in[pair];
    const bool resume = ti.i > 0;
    if (resume) {
        const short2 *src = a.tie_pool + ti.off;
        for (int q = WARP ? (int)(threadIdx.x & 31) : 0; q < ti.len; q += WARP ? 32 : 1) path[q] = src[q];
        if (WARP) __syncwarp();
        i = ti.i; j = ti.j; len = ti.len; c = ti.c;      // (j == m: the window's last column)
    }
    int ts_i = 0, ts_j = 0, ts_len = 0, ts_c = 0;        // TIE3: where the walk first met a marked cell
    bool ts_seen = false;
    if (i & ISTAR_TIE) { tie = 1; i &= ~ISTAR_TIE; }
    if (WARP && (threadIdx.x & 31) != 0) {
        // the other lanes of the pair's warp wait for lane 0's walk (the shuffles below)
    } else if (i <= 0) {
        st |= 2;                      // CRT_ST_NO_POSITIVE
    } else {
        seek(i, j);
        {   // first chunk: also fetch the row above it (the general prefetch reaches two rows up)
            const int t = w_trow + w_l;
            const int idx = (w_strip * u.tchunks + (t >> 2)) * 32 + w_l;
            if ((t >> 2) >= 1) { prefetch_tb(tbu + idx - 32); if (w_l > 0) prefetch_tb(tbu + idx - 33); }
        }
        if (!resume)
            while (j > 1 && (code() & 1u) == 0u) { --j; col_left(); }      // first column of row i* that attains the maximum
        if (tie) ts_seen = true;      // in doubt before the first move: no window
        while (i > 0 && j > 0) {
            if (zreg && is_zero_cell(i, j)) break;
            // one instruction stream for the three moves (the threads of a warp walk different pairs): diag = H == diag + S,
            // left = H == left (and not diag), up otherwise; priority diag > left > up (dynamic_time_warping.py:260-277)
            const unsigned cd = code();
            if (TIE3 && tie && !ts_seen) { ts_seen = true; ts_i = i; ts_j = j; ts_len = len; ts_c = c; }
            const bool diag = (cd & 2u) == 0u, left = !diag && (cd & 1u) == 0u;
            const bool di = diag || !left, dj = diag || left;
            i -= di ? 1 : 0; j -= dj ? 1 : 0; w_trow -= di ? 1 : 0;
            if (dj) col_left();
            path[len++] = make_short2(di ? (short)i : (short)-1, dj ? (short)j : (short)-1);
            c += diag ? 1 : 0;
        }
    }
    if (TIE3 && tie && a.tie_out && (!WARP || (threadIdx.x & 31) == 0)) {
        // save the prefix of a marked pair for the windowed re-run (0.5 % of the pairs of C3)
        TieInfo t{0, 0, 0, 0, 0};
        if (ts_i > 0 && !zreg) {
            const unsigned long long off = atomicAdd(a.tie_pool_used, (unsigned long long)ts_len);
            if (off + (unsigned long long)ts_len <= a.tie_pool_cap) {
                for (int q = 0; q < ts_len; ++q) a.tie_pool[off + q] = path[q];
                t = TieInfo{ts_i, ts_j, ts_len, ts_c, (long long)off};
            }
        }
        a.tie_out[pair] = t;
    }
    if (WARP) {
        __syncwarp();                 // lane 0's path stores are visible to the warp
        len = __shfl_sync(FULL, len, 0); c = __shfl_sync(FULL, c, 0); st = __shfl_sync(FULL, st, 0); tie = __shfl_sync(FULL, tie, 0);
        trace_tail_warp(a, pair, path, len, c, st, tie, n, m_full, A, B, ceni, cenj);
    } else {
        trace_tail(a, pair, path, len, c, st, tie, n, m_full, A, B, ceni, cenj);
    }
}

template <int C, bool TIE3>
__global__ void __launch_bounds__(TRACE_THREADS, CRT_TRACE_MINB) k_trace(TraceArgs a, int n_units, int n_dense)
{
    trace_pair<C, TIE3, false>(a, n_units, n_dense);
}
template <int C, bool TIE3>
__global__ void __launch_bounds__(TRACE_THREADS, 4) k_trace_w(TraceArgs a, int n_units, int n_dense)
{
    trace_pair<C, TIE3, true>(a, n_units, n_dense);
}

// Stage-2 row records, one thread per row of the unit's stream: (x - m1) R^T + m2  == the reference's frame up
// to a rigid motion of BOTH chains (superposition_functions.py:57-58 rotates chain 2 instead; the Gaussian only
// sees distances).  fp32: relative to chain j's centroid and scaled by sqrt(gamma_c log2 e).
__global__ void __launch_bounds__(256) k_rows2(TraceArgs a, int n_units)
{
    if ((int)blockIdx.x >= n_units) return;
    const Unit u = a.units[blockIdx.x];
    const double *cj = a.centroid + (long long)u.col_chain * 3;
    const double c0 = cj[0], c1 = cj[1], c2 = cj[2];
    for (int g = threadIdx.x; g < u.G; g += blockDim.x) {
        const long long r = u.row_base + g;
        const int meta = a.meta[r];
        const double *xf = a.xform + (long long)(u.pair_base + (meta >> 2) - u.row_chain0) * XF;
        const double ax = a.coords[r * 3], ay = a.coords[r * 3 + 1], az = a.coords[r * 3 + 2];
        double y0 = ax, y1 = ay, y2 = az;
        if (xf[15] != 0.0) {
            const double x0 = ax - xf[9], x1 = ay - xf[10], x2 = az - xf[11];
            y0 = (x0 * xf[0] + x1 * xf[1] + x2 * xf[2]) + xf[12];
            y1 = (x0 * xf[3] + x1 * xf[4] + x2 * xf[5]) + xf[13];
            y2 = (x0 * xf[6] + x1 * xf[7] + x2 * xf[8]) + xf[14];
        }
        const long long idx = u.rows2_base + g;
        if (a.rows2_f32) {
            float4 v;
            v.x = (float)((y0 - c0) * (double)a.scale2);
            v.y = (float)((y1 - c1) * (double)a.scale2);
            v.z = (float)((y2 - c2) * (double)a.scale2);
            v.w = __int_as_float(meta);
            reinterpret_cast<float4 *>(a.rows2)[idx] = v;
        } else {
            double *o = reinterpret_cast<double *>(a.rows2) + idx * 4;
            o[0] = y0; o[1] = y1; o[2] = y2; o[3] = __longlong_as_double((long long)meta);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// Preprocessing of the uploaded chains (device side of crt_set_chains).
// ------------------------------------------------------------------------------------------------------------
struct PrepArgs {
    const double *coords, *tensors;
    const long long *offsets;
    int *chain_of;            // [sumL], written by k_chain_of
    int n_chains, d;
    long long total;
    const double *mean;       // [32] global tensor mean (k_tensor_stats / k_tensor_mean)
    double g2;                // gamma_tensor * log2(e)
    double scale2;            // sqrt(gamma_coords * log2(e))
    float *rec32; int rs32, d32;
    double *rec64; int d64;
    int *meta;
    float4 *cols2;
    double *centroid;
};

// chain index of every residue (bisection over the offsets), one thread per residue
__global__ void k_chain_of(PrepArgs a)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.total) return;
    int lo = 0, hi = a.n_chains - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (a.offsets[mid] <= r) lo = mid; else hi = mid - 1;
    }
    a.chain_of[r] = lo;
}

// Global tensor mean and the finiteness check of the uploaded arrays, deterministic for a given input: block b sums a
// fixed contiguous slice in a fixed order (thread-strided sums, then a fixed shared-memory tree), k_tensor_mean adds the
// block partials in block order.  partial: [STATS_BLOCKS][32], flag: nonzero if any coordinate/tensor value is not finite.
constexpr int STATS_BLOCKS = 128, STATS_THREADS = 256;
__global__ void __launch_bounds__(STATS_THREADS) k_tensor_stats(const double *tensors, const double *coords, long long total, int d,
                                                                 double *partial, int *flag)
{
    __shared__ double sh[STATS_THREADS];
    const long long per = (total + STATS_BLOCKS - 1) / STATS_BLOCKS;
    const long long r0 = (long long)blockIdx.x * per, r1 = min(total, r0 + per);
    bool bad = false;
    for (int k = 0; k < d; ++k) {
        double acc = 0.0;
        for (long long r = r0 + threadIdx.x; r < r1; r += STATS_THREADS) {
            const double v = tensors[r * d + k];
            acc += v;
            bad |= !isfinite(v);
        }
        sh[threadIdx.x] = acc;
        __syncthreads();
        for (int w = STATS_THREADS / 2; w > 0; w >>= 1) {
            if ((int)threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
            __syncthreads();
        }
        if (threadIdx.x == 0) partial[blockIdx.x * 32 + k] = sh[0];
        __syncthreads();
    }
    for (long long q = r0 * 3 + threadIdx.x; q < r1 * 3; q += STATS_THREADS) bad |= !isfinite(coords[q]);
    if (bad) atomicOr(flag, 1);
}

__global__ void k_tensor_mean(const double *partial, long long total, int d, double *mean)
{
    const int k = threadIdx.x;
    if (k >= 32) return;
    double acc = 0.0;
    if (k < d)
        for (int b = 0; b < STATS_BLOCKS; ++b) acc += partial[b * 32 + k];
    mean[k] = k < d ? acc / (double)total : 0.0;
}

// One warp per chain: the warp stages tiles of 96 residues in shared memory (coalesced loads, the next tile in flight while this
// one is summed) and lanes 0..2 each add one coordinate left to right: the same order of float64 adds as the serial loop of one
// thread per chain this replaces (so the same bits), without its chain of un-overlapped load round trips (60 us for a
// 375-residue node of a tree level).
constexpr int CEN_TILE = 96;
__global__ void __launch_bounds__(128) k_centroid(PrepArgs a)
{
    __shared__ double tile[4][CEN_TILE * 3];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ch = blockIdx.x * 4 + warp;
    if (ch >= a.n_chains) return;
    const long long b = a.offsets[ch], e = a.offsets[ch + 1];
    const long long n3 = (e - b) * 3;
    const double *src = a.coords + b * 3;
    double acc = 0.0;
    double nxt[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) { const long long x = (long long)q * 32 + lane; nxt[q] = x < n3 ? src[x] : 0.0; }
    for (long long t0 = 0; t0 < n3; t0 += CEN_TILE * 3) {
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 9; ++q) tile[warp][q * 32 + lane] = nxt[q];
        __syncwarp();
        const long long t1 = t0 + CEN_TILE * 3;
        if (t1 < n3) {
#pragma unroll
            for (int q = 0; q < 9; ++q) { const long long x = t1 + (long long)q * 32 + lane; nxt[q] = x < n3 ? src[x] : 0.0; }
        }
        if (lane < 3) {
            const int rows = (int)((n3 - t0 < CEN_TILE * 3 ? n3 - t0 : CEN_TILE * 3) / 3);
            for (int r = 0; r < rows; ++r) acc += tile[warp][r * 3 + lane];
        }
    }
    const double inv = e > b ? 1.0 / (double)(e - b) : 0.0;
    if (lane < 3) a.centroid[(long long)ch * 3 + lane] = acc * inv;
}

__global__ void k_prep(PrepArgs a)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.total) return;
    const int ch = a.chain_of[r];
    const bool first = r == a.offsets[ch], last = r + 1 == a.offsets[ch + 1];
    const int meta = make_meta(ch, first, last);
    a.meta[r] = meta;
    const double *t = a.tensors + r * a.d;
    const double sc = sqrt(2.0 * a.g2);
    double nn = 0.0;
    float *o = a.rec32 + r * a.rs32;
    for (int k = 0; k < a.d32; ++k) {
        const double x = k < a.d ? t[k] - a.mean[k] : 0.0;
        nn += x * x;
        o[k] = (float)(sc * x);
    }
    o[a.d32] = (float)(-a.g2 * nn);
    o[a.d32 + 1] = 1.0f;
    for (int k = a.d32 + 2; k < a.rs32; ++k) o[k] = 0.f;
    double *o64 = a.rec64 + r * a.d64;
    for (int k = 0; k < a.d64; ++k) o64[k] = k < a.d ? t[k] : 0.0;
    const double *cen = a.centroid + (long long)ch * 3;
    float4 v;
    v.x = (float)((a.coords[r * 3] - cen[0]) * a.scale2);
    v.y = (float)((a.coords[r * 3 + 1] - cen[1]) * a.scale2);
    v.z = (float)((a.coords[r * 3 + 2] - cen[2]) * a.scale2);
    v.w = 0.f;
    a.cols2[r] = v;
}

// FP32 FFMA peak probe: 8 independent chains per thread, 148 * k CTAs.
__global__ void k_ffma_peak(float *out, int iters)
{
    float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
    const float a = 0.999f, b = 1e-3f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            x0 = __fmaf_rn(x0, a, b); x1 = __fmaf_rn(x1, a, b); x2 = __fmaf_rn(x2, a, b); x3 = __fmaf_rn(x3, a, b);
            x4 = __fmaf_rn(x4, a, b); x5 = __fmaf_rn(x5, a, b); x6 = __fmaf_rn(x6, a, b); x7 = __fmaf_rn(x7, a, b);
        }
    }
    if (x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 == 12345.678f) out[0] = x0;
}

}  // namespace crt
