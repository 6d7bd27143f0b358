// crt_consumers.cuh -- what consumes the multiple alignment, on the device:
//
//   k_aln_bits / k_coverage_gap   make_coverage_gap_distance_matrix            (multiple_alignment.py:45-56)
//   k_core_mask                   core columns of superpose / superpose_core    (multiple_alignment.py:854-880)
//   k_superpose                   superpose_core / superpose_reference / superpose_references: Kabsch on the chosen
//                                 columns + apply_rotran on the whole chain      (multiple_alignment.py:869-950,
//                                                                                 superposition_functions.py:6-35, 63-80)
//   k_fmt_rowlen / k_fmt_scan /   helper.write_distance_matrix: "%.4f" text, byte-compatible with Python's
//   k_fmt_write                   f"{x:.4f}" (correctly rounded, ties to even)  (helper.py:183-203)
//   k_fasta                       MultipleAlignment.write_alignment / to_sequence_alignment (multiple_alignment.py:287-309)
//
// All of it is integer / byte / float64 work bound by HBM or by latency; nothing here is a contraction.
#pragma once
#include "crt_kernels.cuh"

namespace crt {

// ------------------------------------------------------------------------------------------------------------
// Presence bit masks of the alignment, transposed: bitsT[w * N + p] bit b  <=>  aln[p][32 w + b] != -1.
// One warp per (protein, word): 32 consecutive columns are one coalesced 256-byte load and one ballot.
// present[p] = number of residues of protein p in the alignment.  bad: set when a value is < -1 or >= the chain length.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_aln_bits(const long long *aln, int N, long long A, int W, const long long *offsets,
                                                  unsigned *bitsT, int *present, int *bad)
{
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= (long long)N * W) return;
    const int p = (int)(warp / W), w = (int)(warp - (long long)p * W);
    const long long q = (long long)w * 32 + lane;
    long long v = -1;
    if (q < A) v = aln[(long long)p * A + q];
    if (v < -1 || (offsets && v >= offsets[p + 1] - offsets[p])) atomicOr(bad, 1);
    const unsigned word = __ballot_sync(FULL, v >= 0);
    if (lane == 0) {
        bitsT[(long long)w * N + p] = word;
        if (word) atomicAdd(present + p, __popc(word));
    }
}

// distance[i][j] = (# columns where i has a residue and j a gap) / length_i, aligning[i][j] = length_i - that count.
// grid (ceil(N / 256), N): block row i keeps its words in shared memory, thread j reads the transposed masks coalesced.
__global__ void __launch_bounds__(256) k_coverage_gap(const unsigned *bitsT, const int *present, int N, int W, double *distance,
                                                      int *aligning)
{
    extern __shared__ unsigned wi[];
    const int i = blockIdx.y;
    for (int w = threadIdx.x; w < W; w += blockDim.x) wi[w] = bitsT[(long long)w * N + i];
    __syncthreads();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    int gaps = 0;
    for (int w = 0; w < W; ++w) gaps += __popc(wi[w] & ~bitsT[(long long)w * N + j]);
    const int len = present[i];
    distance[(long long)i * N + j] = (double)gaps / (double)len;       // int64 / int64 true division, :54
    aligning[(long long)i * N + j] = len - gaps;
}

// core[w] = AND over all proteins of their presence words (columns where nobody has a gap, :856-862)
__global__ void __launch_bounds__(256) k_core_mask(const unsigned *bitsT, int N, int W, unsigned *core)
{
    __shared__ unsigned part[8];
    const int w = blockIdx.x;
    unsigned acc = 0xffffffffu;
    for (int p = threadIdx.x; p < N; p += blockDim.x) acc &= bitsT[(long long)w * N + p];
    acc = __reduce_and_sync(FULL, acc);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int q = 1; q < (int)(blockDim.x >> 5); ++q) acc &= part[q];
        core[w] = acc;
    }
}

// ------------------------------------------------------------------------------------------------------------
// One warp per (reference, member) superposition.
//   columns: cols != nullptr -> the given column list (core mode), else every column where both have a residue
//            (helper.get_common_positions, helper.py:12-42).
//   core mode (center_ref): the reference's columns are first centred on their own centroid (ref_coords -= ref_centroid,
//            :884-885), a member equal to the reference is only translated by -ref_centroid (:887-888).
//   R, t = paired_svd_superpose(X1, X2) (superposition_functions.py:6-35); dst[member] = src[member] R + t (apply_rotran).
//   fewer than min_common usable columns (4 where the reference asserts len(pos_1) > 3, :918 / :941; 1 in core mode):
//            nothing is written, ncommon tells the host.
// src and dst may alias when no member of the launch is the reference of another pair of the same launch.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

__global__ void __launch_bounds__(128) k_superpose(const double *src, double *dst, const long long *offsets, const long long *aln,
                                                   long long A, const int *ref, const int *mem, int n_pairs, const int *cols,
                                                   int n_cols, int center_ref, int min_common, double *rot, double *tran, int *ncommon)
{
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n_pairs) return;
    const int r = ref[w], p = mem[w];
    const long long *ar = aln + (long long)r * A, *ap = aln + (long long)p * A;
    const double *Xr = src + offsets[r] * 3, *Xp = src + offsets[p] * 3;
    double *Yp = dst + offsets[p] * 3;
    const int Lp = (int)(offsets[p + 1] - offsets[p]);
    const long long nc = cols ? n_cols : A;
    // pass 1: usable columns and raw sums
    int c = 0;
    double s1[3] = {0, 0, 0}, s2[3] = {0, 0, 0};
    for (long long k = lane; k < nc; k += 32) {
        const long long q = cols ? cols[k] : k;
        const long long a = ar[q], b = ap[q];
        if (a < 0 || b < 0) continue;
        ++c;
#pragma unroll
        for (int d = 0; d < 3; ++d) { s1[d] += Xr[a * 3 + d]; s2[d] += Xp[b * 3 + d]; }
    }
    c = __reduce_add_sync(FULL, c);
#pragma unroll
    for (int d = 0; d < 3; ++d) { s1[d] = warp_sum(s1[d]); s2[d] = warp_sum(s2[d]); }
    if (lane == 0) ncommon[w] = c;
    double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, t[3] = {0, 0, 0};
    if (center_ref && r == p) {                           // the reference of superpose_core: coordinates -= ref_centroid
        if (c == 0) return;
#pragma unroll
        for (int d = 0; d < 3; ++d) t[d] = -(s1[d] / (double)c);
    } else {
        if (c < min_common) return;
        double pre[3] = {0, 0, 0}, m1[3], m2[3];
        if (center_ref) {
            // X1 = ref_core - ref_centroid, then paired_svd_superpose takes its mean again (it is ~1e-15, not 0)
#pragma unroll
            for (int d = 0; d < 3; ++d) pre[d] = s1[d] / (double)c;
            double z[3] = {0, 0, 0};
            for (long long k = lane; k < nc; k += 32) {
                const long long q = cols ? cols[k] : k;
                const long long a = ar[q], b = ap[q];
                if (a < 0 || b < 0) continue;
#pragma unroll
                for (int d = 0; d < 3; ++d) z[d] += __dsub_rn(Xr[a * 3 + d], pre[d]);
            }
#pragma unroll
            for (int d = 0; d < 3; ++d) m1[d] = warp_sum(z[d]) / (double)c;
        } else {
#pragma unroll
            for (int d = 0; d < 3; ++d) m1[d] = s1[d] / (double)c;
        }
#pragma unroll
        for (int d = 0; d < 3; ++d) m2[d] = s2[d] / (double)c;
        // pass 2: correlation matrix C = X2c^T X1c
        double Cm[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (long long k = lane; k < nc; k += 32) {
            const long long q = cols ? cols[k] : k;
            const long long a = ar[q], b = ap[q];
            if (a < 0 || b < 0) continue;
            double x1[3], x2[3];
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                x1[d] = __dsub_rn(__dsub_rn(Xr[a * 3 + d], pre[d]), m1[d]);
                x2[d] = __dsub_rn(Xp[b * 3 + d], m2[d]);
            }
#pragma unroll
            for (int u = 0; u < 3; ++u)
#pragma unroll
                for (int v = 0; v < 3; ++v) Cm[u * 3 + v] += x2[u] * x1[v];
        }
#pragma unroll
        for (int q = 0; q < 9; ++q) Cm[q] = warp_sum(Cm[q]);
        if (lane == 0) kabsch_rotation(Cm, R);
#pragma unroll
        for (int q = 0; q < 9; ++q) R[q] = __shfl_sync(FULL, R[q], 0);
#pragma unroll
        for (int b = 0; b < 3; ++b)
            t[b] = __dsub_rn(m1[b], __dadd_rn(__dadd_rn(__dmul_rn(m2[0], R[b]), __dmul_rn(m2[1], R[3 + b])), __dmul_rn(m2[2], R[6 + b])));
    }
    if (lane == 0) {
        for (int q = 0; q < 9; ++q) rot[(long long)w * 9 + q] = R[q];
        for (int q = 0; q < 3; ++q) tran[(long long)w * 3 + q] = t[q];
    }
    __syncwarp();
    const bool translate_only = center_ref && r == p;
    for (int res = lane; res < Lp; res += 32) {
        const double x0 = Xp[res * 3], x1 = Xp[res * 3 + 1], x2 = Xp[res * 3 + 2];
        double y[3];
        if (translate_only) {
            y[0] = __dadd_rn(x0, t[0]); y[1] = __dadd_rn(x1, t[1]); y[2] = __dadd_rn(x2, t[2]);       // x - centroid
        } else {
#pragma unroll
            for (int b = 0; b < 3; ++b)
                y[b] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(x0, R[b]), __dmul_rn(x1, R[3 + b])), __dmul_rn(x2, R[6 + b])), t[b]);
        }
        Yp[res * 3] = y[0]; Yp[res * 3 + 1] = y[1]; Yp[res * 3 + 2] = y[2];
    }
}

// ------------------------------------------------------------------------------------------------------------
// "%.4f" exactly like Python's f"{x:.4f}" / C printf: the decimal expansion of the binary value, rounded to four
// places, ties to even.  |x| < 2^63: integer part by truncation (exact), fraction f = |x| - int (exact), f * 10^4 as an
// exact double-double (product + fma error term); the rounding decision needs only floor(p), p - floor(p) (exact) and
// the sign of the error term, because p - floor(p) and 0.5 are both multiples of ulp(p) while |err| <= ulp(p) / 2.
// |x| >= 2^63: an exact integer M * 2^e, expanded in base 10^9 limbs (rare; distance matrices never get there).
// ------------------------------------------------------------------------------------------------------------
struct Fmt4 {
    unsigned long long ip;     // integer part
    unsigned frac;             // 0..9999
    int neg;                   // sign bit (also for -0.0 and values that round to zero: "-0.0000", like Python)
    int kind;                  // 0 finite < 2^63, 1 nan, 2 inf, 3 finite >= 2^63
};

__device__ __forceinline__ Fmt4 fmt4_decode(double x)
{
    Fmt4 f;
    f.neg = __double_as_longlong(x) < 0;
    f.ip = 0; f.frac = 0; f.kind = 0;
    if (x != x) { f.kind = 1; f.neg = 0; return f; }                  // Python prints "nan" whatever the sign bit
    const double ax = fabs(x);
    if (isinf(ax)) { f.kind = 2; return f; }
    if (ax >= 9223372036854775808.0) { f.kind = 3; return f; }
    f.ip = (unsigned long long)ax;
    const double fr = ax - (double)f.ip;
    const double p = __dmul_rn(fr, 1e4), err = __fma_rn(fr, 1e4, -p);
    const double r = floor(p), fc = p - r;
    unsigned q = (unsigned)r;
    if (fc > 0.5 || (fc == 0.5 && (err > 0.0 || (err == 0.0 && (q & 1u))))) ++q;
    if (q >= 10000u) { q -= 10000u; ++f.ip; }
    f.frac = q;
    return f;
}

__device__ __forceinline__ int u64_digits(unsigned long long v)
{
    if (v < 100000ull) {                                   // the usual case (a distance, a score): comparisons only
        const unsigned u = (unsigned)v;
        return 1 + (u >= 10u) + (u >= 100u) + (u >= 1000u) + (u >= 10000u);
    }
    int n = 5;
    v /= 100000ull;
    while (v) { v /= 10ull; ++n; }
    return n;
}

// expands |x| >= 2^63 (an integer) into base-1e9 limbs; returns the number of limbs
__device__ __noinline__ int fmt4_big_limbs(double ax, unsigned *L /* [36] */)
{
    const unsigned long long bits = (unsigned long long)__double_as_longlong(ax);
    unsigned long long M = (bits & 0xfffffffffffffull) | 0x10000000000000ull;
    int e = (int)((bits >> 52) & 0x7ff) - 1075;                         // ax = M * 2^e, e >= 11 here
    int n = 0;
    while (M) { L[n++] = (unsigned)(M % 1000000000ull); M /= 1000000000ull; }
    while (e > 0) {
        const int k = e > 29 ? 29 : e;
        unsigned long long carry = 0;
        for (int q = 0; q < n; ++q) {
            const unsigned long long v = ((unsigned long long)L[q] << k) + carry;
            L[q] = (unsigned)(v % 1000000000ull);
            carry = v / 1000000000ull;
        }
        while (carry) { L[n++] = (unsigned)(carry % 1000000000ull); carry /= 1000000000ull; }
        e -= k;
    }
    return n;
}

__device__ __noinline__ int fmt4_big(double x, char *o /* nullable: length only */)
{
    unsigned L[36];
    const int n = fmt4_big_limbs(fabs(x), L);
    const int neg = __double_as_longlong(x) < 0;
    const int top = u64_digits(L[n - 1]);
    const int len = neg + top + 9 * (n - 1) + 5;
    if (!o) return len;
    int k = 0;
    if (neg) o[k++] = '-';
    for (int q = n - 1; q >= 0; --q) {
        const int nd = q == n - 1 ? top : 9;
        unsigned v = L[q];
        for (int z = nd - 1; z >= 0; --z) { o[k + z] = (char)('0' + v % 10u); v /= 10u; }
        k += nd;
    }
    o[k++] = '.'; o[k++] = '0'; o[k++] = '0'; o[k++] = '0'; o[k++] = '0';
    return len;
}

__device__ __forceinline__ int fmt4_len(const Fmt4 &f, double x)
{
    if (f.kind == 1) return 3;
    if (f.kind == 2) return 3 + f.neg;
    if (f.kind == 3) return fmt4_big(x, nullptr);
    return f.neg + u64_digits(f.ip) + 5;
}

// kinds 0..2 into o (at most 25 characters); returns the length
__device__ __forceinline__ int fmt4_write(const Fmt4 &f, char *o)
{
    int k = 0;
    if (f.kind == 1) { o[0] = 'n'; o[1] = 'a'; o[2] = 'n'; return 3; }
    if (f.neg) o[k++] = '-';
    if (f.kind == 2) { o[k] = 'i'; o[k + 1] = 'n'; o[k + 2] = 'f'; return k + 3; }
    const int nd = u64_digits(f.ip);
    if (f.ip < 4294967296ull) {                              // 32-bit digit loop (division by a constant = multiply + shift)
        unsigned v = (unsigned)f.ip;
        for (int z = nd - 1; z >= 0; --z) { o[k + z] = (char)('0' + v % 10u); v /= 10u; }
    } else {
        unsigned long long v = f.ip;
        for (int z = nd - 1; z >= 0; --z) { o[k + z] = (char)('0' + (int)(v % 10ull)); v /= 10ull; }
    }
    k += nd;
    unsigned q = f.frac;
    o[k] = '.';
    o[k + 4] = (char)('0' + q % 10u); q /= 10u;
    o[k + 3] = (char)('0' + q % 10u); q /= 10u;
    o[k + 2] = (char)('0' + q % 10u); q /= 10u;
    o[k + 1] = (char)('0' + q);
    return k + 5;
}

constexpr int FMT_THREADS = 256;
constexpr int FMT_STAGE = FMT_THREADS * 12 + 32;      // bytes of text staged per tile of 256 values (typical: 7 per value)

__device__ __forceinline__ int block_sum_256(int v, int *sh /* [8] */)
{
    v = __reduce_add_sync(FULL, v);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    int tot = 0;
#pragma unroll
    for (int q = 0; q < FMT_THREADS / 32; ++q) tot += sh[q];
    __syncthreads();
    return tot;
}

// bytes of line i: name + ' ' + values joined by ' ' + '\n' (helper.py:200-203)
__global__ void __launch_bounds__(FMT_THREADS) k_fmt_rowlen(const double *M, int C, const long long *name_off, long long *row_len)
{
    __shared__ int sh[8];
    const int i = blockIdx.x;
    long long tot = 0;
    for (int j0 = 0; j0 < C; j0 += FMT_THREADS) {
        const int j = j0 + threadIdx.x;
        int len = 0;
        if (j < C) { const double x = M[(long long)i * C + j]; len = fmt4_len(fmt4_decode(x), x) + 1; }
        tot += block_sum_256(len, sh);
    }
    if (threadIdx.x == 0) row_len[i] = (name_off[i + 1] - name_off[i]) + 1 + (C > 0 ? tot : 1);
}

// exclusive scan of n values by one block (n = number of lines; a few thousand): off[i] = first + sum_{k<i} len[k], off[n] = total
__global__ void __launch_bounds__(1024) k_fmt_scan(const long long *len, int n, long long first, long long *off)
{
    __shared__ long long part[1024];
    const int per = (n + 1023) / 1024, t = threadIdx.x;
    const int lo = min(t * per, n), hi = min(lo + per, n);
    long long s = 0;
    for (int q = lo; q < hi; ++q) s += len[q];
    part[t] = s;
    __syncthreads();
    if (t == 0) {
        long long run = first;
        for (int q = 0; q < 1024; ++q) { const long long v = part[q]; part[q] = run; run += v; }
        off[n] = run;
    }
    __syncthreads();
    long long run = part[t];
    for (int q = lo; q < hi; ++q) { off[q] = run; run += len[q]; }
}

// One block per line.  Tiles of 256 values: format into registers, block scan of the lengths, text staged in shared memory at
// the same 16-byte phase as its destination, then written with aligned 128-bit stores (partial head / tail chunks bytewise).
__global__ void __launch_bounds__(FMT_THREADS) k_fmt_write(const double *M, int C, const char *names, const long long *name_off,
                                                           const long long *row_off, char *out)
{
    __shared__ __align__(16) char stage[FMT_STAGE];
    __shared__ int wsum[FMT_THREADS / 32];
    const int i = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    char *dst = out + row_off[i];
    const long long n0 = name_off[i], nl = name_off[i + 1] - n0;
    for (long long k = tid; k < nl; k += FMT_THREADS) dst[k] = names[n0 + k];
    if (tid == 0) { dst[nl] = ' '; if (C == 0) dst[nl + 1] = '\n'; }
    long long base = nl + 1;
    for (int j0 = 0; j0 < C; j0 += FMT_THREADS) {
        const int j = j0 + tid;
        double x = 0.0;
        Fmt4 f;
        f.kind = 0;
        int len = 0;
        if (j < C) { x = M[(long long)i * C + j]; f = fmt4_decode(x); len = fmt4_len(f, x) + 1; }
        // block exclusive scan of len
        int inc = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += v; }
        if (lane == 31) wsum[wid] = inc;
        __syncthreads();
        int woff = 0, total = 0;
#pragma unroll
        for (int q = 0; q < FMT_THREADS / 32; ++q) { const int v = wsum[q]; if (q < wid) woff += v; total += v; }
        const int off = woff + inc - len;
        const int big = __syncthreads_or(j < C && f.kind == 3);
        const char sep = (j == C - 1) ? '\n' : ' ';
        const int shift = (int)((unsigned long long)(dst + base) & 15ull);
        if (!big && shift + total <= FMT_STAGE) {
            if (j < C) {                                                // digits go straight into the staging tile (no local buffer)
                char *s = stage + shift + off;
                const int n = fmt4_write(f, s);
                s[n] = sep;
            }
            __syncthreads();
            char *g = dst + base - shift;                               // 16-byte aligned
            const int end = shift + total;
            for (int lo = tid * 16; lo < end; lo += FMT_THREADS * 16) {
                const int hi = lo + 16;
                if (lo >= shift && hi <= end) {
                    *reinterpret_cast<uint4 *>(g + lo) = *reinterpret_cast<const uint4 *>(stage + lo);
                } else {
                    const int a = lo > shift ? lo : shift, b = hi < end ? hi : end;
                    for (int k = a; k < b; ++k) g[k] = stage[k];
                }
            }
            __syncthreads();
        } else if (j < C) {                                             // a tile with huge numbers: straight to global memory
            char *o = dst + base + off;
            const int n = f.kind == 3 ? fmt4_big(x, o) : fmt4_write(f, o);
            o[n] = sep;
        }
        base += total;
    }
}

// ------------------------------------------------------------------------------------------------------------
// FASTA text of the alignment: for every protein ">name\n" + (sequence[aln[k]] or '-') for k < A + "\n"  (:299-309).
// One block per protein; rec_off[p] = byte offset of its record.  bad: an index outside the sequence.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fasta(const long long *aln, long long A, const char *seqs, const long long *seq_off,
                                               const char *names, const long long *name_off, const long long *rec_off, char *out,
                                               int *bad)
{
    const int p = blockIdx.x;
    char *dst = out + rec_off[p];
    const long long n0 = name_off[p], nl = name_off[p + 1] - n0, s0 = seq_off[p], sl = seq_off[p + 1] - s0;
    if (threadIdx.x == 0) { dst[0] = '>'; dst[1 + nl] = '\n'; dst[2 + nl + A] = '\n'; }
    for (long long k = threadIdx.x; k < nl; k += blockDim.x) dst[1 + k] = names[n0 + k];
    char *body = dst + 2 + nl;
    const long long *ap = aln + (long long)p * A;
    for (long long k = threadIdx.x; k < A; k += blockDim.x) {
        const long long v = ap[k];
        char ch = '-';
        if (v >= 0) {
            if (v < sl) ch = seqs[s0 + v];
            else { atomicOr(bad, 1); ch = '?'; }
        } else if (v < -1) atomicOr(bad, 1);
        body[k] = ch;
    }
}


// ------------------------------------------------------------------------------------------------------------
// Fast-mode guide matrix (align_from_structure_files with full=False, multiple_alignment.py:503-511):
//   k_count_matrix   make_count_matrix (:128-134): out[i][r] += 1 for every shapemer index r of protein i
//   k_braycurtis     braycurtis (:137-145): out[i][j] = sum_k |a_ik - b_jk| / sum_k |a_ik + b_jk|, float64, both sums taken
//                    in k order with separate roundings like numba's sequential array.sum() (for count data every partial sum
//                    is an exact integer anyway).
// 32 x 32 pairs per block of 256 threads (2 x 2 pairs per thread), k in chunks of 32 staged in shared memory.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_count_matrix(const long long *idx, const long long *off, int N, int K, double *out, int *bad)
{
    const int p = blockIdx.x;
    for (long long q = off[p] + threadIdx.x; q < off[p + 1]; q += blockDim.x) {
        const long long r = idx[q];
        if (r < 0 || r >= K) { atomicOr(bad, 1); continue; }
        atomicAdd(out + (long long)p * K + r, 1.0);          // counts stay far below 2^53: exact in any order
    }
}

constexpr int BC_TILE = 32, BC_K = 32;

__global__ void __launch_bounds__(256) k_braycurtis(const double *A, int n1, const double *Bm, int n2, int K, double *out)
{
    __shared__ double sa[BC_TILE][BC_K + 1], sb[BC_TILE][BC_K + 1];
    const int i0 = blockIdx.y * BC_TILE, j0 = blockIdx.x * BC_TILE;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;          // thread -> pairs (ty, ty + 16) x (tx, tx + 16)
    double num[2][2] = {{0, 0}, {0, 0}}, den[2][2] = {{0, 0}, {0, 0}};
    for (int k0 = 0; k0 < K; k0 += BC_K) {
        for (int e = threadIdx.x; e < BC_TILE * BC_K; e += 256) {
            const int r = e / BC_K, k = e % BC_K;
            sa[r][k] = (i0 + r < n1 && k0 + k < K) ? A[(long long)(i0 + r) * K + k0 + k] : 0.0;
            sb[r][k] = (j0 + r < n2 && k0 + k < K) ? Bm[(long long)(j0 + r) * K + k0 + k] : 0.0;
        }
        __syncthreads();
        const int kn = min(BC_K, K - k0);
        for (int k = 0; k < kn; ++k) {
            const double a0 = sa[ty][k], a1 = sa[ty + 16][k], b0 = sb[tx][k], b1 = sb[tx + 16][k];
            num[0][0] = __dadd_rn(num[0][0], fabs(__dsub_rn(a0, b0))); den[0][0] = __dadd_rn(den[0][0], fabs(__dadd_rn(a0, b0)));
            num[0][1] = __dadd_rn(num[0][1], fabs(__dsub_rn(a0, b1))); den[0][1] = __dadd_rn(den[0][1], fabs(__dadd_rn(a0, b1)));
            num[1][0] = __dadd_rn(num[1][0], fabs(__dsub_rn(a1, b0))); den[1][0] = __dadd_rn(den[1][0], fabs(__dadd_rn(a1, b0)));
            num[1][1] = __dadd_rn(num[1][1], fabs(__dsub_rn(a1, b1))); den[1][1] = __dadd_rn(den[1][1], fabs(__dadd_rn(a1, b1)));
        }
        __syncthreads();
    }
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int v = 0; v < 2; ++v) {
            const int i = i0 + ty + 16 * u, j = j0 + tx + 16 * v;
            if (i < n1 && j < n2) out[(long long)i * n2 + j] = num[u][v] / den[u][v];
        }
}

}  // namespace crt
