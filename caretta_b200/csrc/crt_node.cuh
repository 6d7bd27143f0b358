// crt_node.cuh -- one node of the progressive alignment on the device, all float64.
//
// Reference: MultipleAlignment.progressive_align / make_intermediate_node, multiple_alignment.py:172-253:
//   score_matrix  = Protein.score_function(n1, n2)                       (:203-205, :321-349; stage 1 = the pair kernels)
//   score_matrix += make_score_matrix(w1 * mult1, w2 * mult2, gaussian, gamma_weight)              (:206-210)
//   aln_1, aln_2  = dtw_align(arange, arange, score_matrix, gap_open, gap_extend)   (:211-214; k_dtw_fill / k_dtw_trace)
//   intermediate  = Protein.mean_function(n1, n2, aln_1, aln_2)          (:215-216, :351-383)
//   weights       = get_mean_weights(w1, w2, aln_1, aln_2)                (:217, :73-82)
#pragma once
#include "crt_kernels.cuh"
#include "crt_dp_batch.cuh"

namespace crt {

// S[a][b] = exp(-gamma_c |c1'[a] - c2'[b]|^2) + exp(-gamma_w (w1[a] mult1 - w2[b] mult2)^2), with the chains in the frame of
// paired_svd_superpose_with_subset (superposition_functions.py:38-60): c1' = c1 - mean(common_1),
// c2' = (c2 - mean(common_2)) R; raw coordinates when the stage-1 alignment has <= 3 common positions (:337-342).
// xf = k_trace's transform record of the pair: R[9], mean_1[3], mean_2[3], superpose flag.
__device__ __forceinline__ void node_score_cell(long long q, const double *c1, int n, const double *c2, int m, const double *xf,
                                                const double *w1, const double *w2, double mult1, double mult2,
                                                double neg_gamma_c, double neg_gamma_w, double *S)
{
    const int a = (int)(q / m), b = (int)(q - (long long)a * m);
    double x0 = c1[a * 3], x1 = c1[a * 3 + 1], x2 = c1[a * 3 + 2];
    double y0 = c2[b * 3], y1 = c2[b * 3 + 1], y2 = c2[b * 3 + 2];
    if (xf[15] != 0.0) {
        x0 = __dsub_rn(x0, xf[9]); x1 = __dsub_rn(x1, xf[10]); x2 = __dsub_rn(x2, xf[11]);
        const double u0 = __dsub_rn(y0, xf[12]), u1 = __dsub_rn(y1, xf[13]), u2 = __dsub_rn(y2, xf[14]);
        y0 = __dadd_rn(__dadd_rn(__dmul_rn(u0, xf[0]), __dmul_rn(u1, xf[3])), __dmul_rn(u2, xf[6]));
        y1 = __dadd_rn(__dadd_rn(__dmul_rn(u0, xf[1]), __dmul_rn(u1, xf[4])), __dmul_rn(u2, xf[7]));
        y2 = __dadd_rn(__dadd_rn(__dmul_rn(u0, xf[2]), __dmul_rn(u1, xf[5])), __dmul_rn(u2, xf[8]));
    }
    // score_functions.py:11: sequential sum of (x - y)^2 with separate roundings, exp((-gamma) * acc)
    double t = __dsub_rn(x0, y0), acc = __dmul_rn(t, t);
    t = __dsub_rn(x1, y1); acc = __dadd_rn(acc, __dmul_rn(t, t));
    t = __dsub_rn(x2, y2); acc = __dadd_rn(acc, __dmul_rn(t, t));
    const double sc = exp(__dmul_rn(neg_gamma_c, acc));
    double sw = 0.0;                              // neg_gamma_w > 0 (gamma_weight < 0): no weight term (two-structure case, :263-275)
    if (!(neg_gamma_w > 0.0)) {
        const double dw = __dsub_rn(__dmul_rn(w1[a], mult1), __dmul_rn(w2[b], mult2));
        sw = exp(__dmul_rn(neg_gamma_w, __dmul_rn(dw, dw)));
    }
    S[q] = __dadd_rn(sc, sw);
}

__global__ void __launch_bounds__(256) k_node_score(const double *c1, int n, const double *c2, int m, const double *xf,
                                                    const double *w1, const double *w2, double mult1, double mult2,
                                                    double neg_gamma_c, double neg_gamma_w, double *S)
{
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= (long long)n * m) return;
    node_score_cell(q, c1, n, c2, m, xf, w1, w2, mult1, mult2, neg_gamma_c, neg_gamma_w, S);
}

// A whole tree level at once (crt_progressive_level): node = blockIdx.y, its two children are consecutive chains of the packed
// level arrays (child 1 at residue pr.aln_off, child 2 at pr.aln_off + n); xform / mult are indexed by node.
__global__ void __launch_bounds__(256) k_level_score(const DpProblem *probs, const double *coords, const double *weights,
                                                     const double *xform, const double *mult, double neg_gamma_c, double neg_gamma_w,
                                                     double *S_all)
{
    const DpProblem pr = probs[blockIdx.y];
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= (long long)pr.n * pr.m) return;
    node_score_cell(q, coords + pr.aln_off * 3, pr.n, coords + (pr.aln_off + pr.n) * 3, pr.m, xform + (long long)blockIdx.y * XF,
                    weights + pr.aln_off, weights + pr.aln_off + pr.n, mult[2 * blockIdx.y], mult[2 * blockIdx.y + 1], neg_gamma_c,
                    neg_gamma_w, S_all + pr.s_off);
}

// flexible=True (multiple_alignment.py:323-326): the score matrix is the Gaussian of the shape tensors alone (no stage-1
// alignment, no superposition), plus the consensus-weight term.  Sequential sum with separate roundings like
// score_functions.py:11.
__global__ void __launch_bounds__(256) k_level_score_flex(const DpProblem *probs, const double *tensors, int d, const double *weights,
                                                          const double *mult, double neg_gamma_t, double neg_gamma_w, double *S_all)
{
    const DpProblem pr = probs[blockIdx.y];
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= (long long)pr.n * pr.m) return;
    const int a = (int)(q / pr.m), b = (int)(q - (long long)a * pr.m);
    const double *x = tensors + (pr.aln_off + a) * d, *y = tensors + (pr.aln_off + pr.n + b) * d;
    double acc = 0.0;
    for (int k = 0; k < d; ++k) {
        const double t = __dsub_rn(x[k], y[k]);
        acc = __dadd_rn(acc, __dmul_rn(t, t));
    }
    const double sc = exp(__dmul_rn(neg_gamma_t, acc));
    double sw = 0.0;
    if (!(neg_gamma_w > 0.0)) {
        const double dw = __dsub_rn(__dmul_rn(weights[pr.aln_off + a], mult[2 * blockIdx.y]),
                                    __dmul_rn(weights[pr.aln_off + pr.n + b], mult[2 * blockIdx.y + 1]));
        sw = exp(__dmul_rn(neg_gamma_w, __dmul_rn(dw, dw)));
    }
    S_all[pr.s_off + q] = __dadd_rn(sc, sw);
}

// Superposition of mean_function (multiple_alignment.py:362-372) over the common positions of the DTW alignment.
// One WARP per node (round 2): lane l takes the alignment columns l, l + 32, ... (two at a time, branch-free: the loads of both
// are in flight together), shuffle-tree sums, lane 0 keeps the 3 x 3 SVD.  One thread summing in alignment order like
// helper.nb_mean_axis_0 (helper.py:45-53) was ~600 dependent round trips twice over (180 us of every tree level); the order of
// the sums moves the rotation by a few ulp, far inside the 1e-9 the goldens hold (the reference's own SVD is LAPACK's).  The
// single-node kernel and the level kernel call the same function, so node-by-node, level and pool paths stay bit-identical.
// xf2: same record layout as xf.  Every lane of the warp must call this.
__device__ inline void node_kabsch_warp(const double *c1, const double *c2, const int *aln1, const int *aln2, int len, double *xf2)
{
    const int lane = threadIdx.x & 31;
    double s1[3] = {0, 0, 0}, s2[3] = {0, 0, 0}, cnt = 0.0;
    for (int q0 = lane; q0 < len; q0 += 64) {
        double v1[2][3], v2[2][3], w[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int q = q0 + 32 * u;
            const int x = q < len ? aln1[q] : -1, y = q < len ? aln2[q] : -1;
            const bool ok = x >= 0 && y >= 0;
            w[u] = ok ? 1.0 : 0.0;
#pragma unroll
            for (int k = 0; k < 3; ++k) { v1[u][k] = c1[(ok ? x : 0) * 3 + k]; v2[u][k] = c2[(ok ? y : 0) * 3 + k]; }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            cnt += w[u];
#pragma unroll
            for (int k = 0; k < 3; ++k) { s1[k] += w[u] * v1[u][k]; s2[k] += w[u] * v2[u][k]; }
        }
    }
    cnt = warp_sum_t(cnt);
#pragma unroll
    for (int k = 0; k < 3; ++k) { s1[k] = warp_sum_t(s1[k]); s2[k] = warp_sum_t(s2[k]); }
    const int c = (int)cnt;
    double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, m1[3] = {0, 0, 0}, m2[3] = {0, 0, 0};
    const bool superpose = c > 3;
    if (superpose) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { m1[k] = s1[k] / (double)c; m2[k] = s2[k] / (double)c; }
        double Cm[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int q0 = lane; q0 < len; q0 += 64) {
            double v1[2][3], v2[2][3], w[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int q = q0 + 32 * u;
                const int x = q < len ? aln1[q] : -1, y = q < len ? aln2[q] : -1;
                const bool ok = x >= 0 && y >= 0;
                w[u] = ok ? 1.0 : 0.0;
#pragma unroll
                for (int k = 0; k < 3; ++k) { v1[u][k] = c1[(ok ? x : 0) * 3 + k]; v2[u][k] = c2[(ok ? y : 0) * 3 + k]; }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int b = 0; b < 3; ++b)
                        Cm[a * 3 + b] += w[u] * ((v2[u][a] - m2[a]) * (v1[u][b] - m1[b]));
        }
#pragma unroll
        for (int k = 0; k < 9; ++k) Cm[k] = warp_sum_t(Cm[k]);
        if (lane == 0) kabsch_rotation(Cm, R);
    }
    if (lane == 0) {
        for (int q = 0; q < 9; ++q) xf2[q] = R[q];
        for (int q = 0; q < 3; ++q) { xf2[9 + q] = m1[q]; xf2[12 + q] = m2[q]; }
        xf2[15] = superpose ? 1.0 : 0.0;
    }
}

__global__ void __launch_bounds__(32) k_node_kabsch(const double *c1, const double *c2, const int *aln1, const int *aln2, const int *len_p, double *xf2)
{
    if (blockIdx.x != 0) return;
    node_kabsch_warp(c1, c2, aln1, aln2, *len_p, xf2);
}

__global__ void __launch_bounds__(32) k_level_kabsch(const DpProblem *probs, int n_nodes, const double *coords, const int *aln1,
                                                     const int *aln2, const int *aln_len, double *xf2)
{
    const int p = blockIdx.x;                 // one warp per node
    if (p >= n_nodes) return;
    const DpProblem pr = probs[p];
    node_kabsch_warp(coords + pr.aln_off * 3, coords + (pr.aln_off + pr.n) * 3, aln1 + pr.aln_off, aln2 + pr.aln_off, aln_len[p],
                     xf2 + (long long)p * XF);
}

// Intermediate node: tensors_mean [len, d], coordinates_mean [len, 3], mean weights [len]; one thread per alignment column.
__device__ __forceinline__ void node_mean_col(int i, const double *t1, const double *c1, const double *w1, const double *t2,
                                              const double *c2, const double *w2, int d, const int *aln1, const int *aln2,
                                              const double *xf2, double *t_out, double *c_out, double *w_out)
{
    const int x = aln1[i], y = aln2[i];
    for (int k = 0; k < d; ++k) {
        double v;
        if (x < 0) v = t2[(size_t)y * d + k];
        else if (y < 0) v = t1[(size_t)x * d + k];
        else v = __dadd_rn(t1[(size_t)x * d + k], t2[(size_t)y * d + k]) / 2;
        t_out[(size_t)i * d + k] = v;
    }
    const bool sp = xf2[15] != 0.0;
    double p[3] = {0, 0, 0}, r[3] = {0, 0, 0};
    if (x >= 0)
        for (int k = 0; k < 3; ++k) p[k] = sp ? __dsub_rn(c1[x * 3 + k], xf2[9 + k]) : c1[x * 3 + k];
    if (y >= 0) {
        if (sp) {
            const double u0 = __dsub_rn(c2[y * 3], xf2[12]), u1 = __dsub_rn(c2[y * 3 + 1], xf2[13]), u2 = __dsub_rn(c2[y * 3 + 2], xf2[14]);
            for (int k = 0; k < 3; ++k)
                r[k] = __dadd_rn(__dadd_rn(__dmul_rn(u0, xf2[k]), __dmul_rn(u1, xf2[3 + k])), __dmul_rn(u2, xf2[6 + k]));
        } else {
            for (int k = 0; k < 3; ++k) r[k] = c2[y * 3 + k];
        }
    }
    for (int k = 0; k < 3; ++k) c_out[i * 3 + k] = x < 0 ? r[k] : (y < 0 ? p[k] : __dadd_rn(p[k], r[k]) / 2);
    double w = 0.0;                                                  // get_mean_weights, multiple_alignment.py:73-82
    if (x >= 0) w = __dadd_rn(w, w1[x]);
    if (y >= 0) w = __dadd_rn(w, w2[y]);
    w_out[i] = w;
}

__global__ void __launch_bounds__(128) k_node_mean(const double *t1, const double *c1, const double *w1, const double *t2,
                                                   const double *c2, const double *w2, int d, const int *aln1, const int *aln2,
                                                   const int *len_p, const double *xf2, double *t_out, double *c_out, double *w_out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *len_p) return;
    node_mean_col(i, t1, c1, w1, t2, c2, w2, d, aln1, aln2, xf2, t_out, c_out, w_out);
}

// get_mean_weights (multiple_alignment.py:73-82) for a given alignment: one thread per column.
__global__ void __launch_bounds__(256) k_mean_weights(const double *w1, const double *w2, const int *aln1, const int *aln2, long long len,
                                                      double *w_out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len) return;
    const int x = aln1[i], y = aln2[i];
    double w = 0.0;
    if (x >= 0) w = __dadd_rn(w, w1[x]);
    if (y >= 0) w = __dadd_rn(w, w2[y]);
    w_out[i] = w;
}

// outputs of node p go to rows pr.aln_off .. pr.aln_off + aln_len[p] of the packed output arrays (capacity n + m rows)
// out_off: row offset of every node's output (nullptr: the node's own rows pr.aln_off of the packed level arrays)
__global__ void __launch_bounds__(128) k_level_mean(const DpProblem *probs, const double *tensors, const double *coords,
                                                    const double *weights, int d, const int *aln1, const int *aln2, const int *aln_len,
                                                    const double *xf2, const long long *out_off, double *t_out, double *c_out, double *w_out)
{
    const DpProblem pr = probs[blockIdx.y];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= aln_len[blockIdx.y]) return;
    const long long a = pr.aln_off, b = pr.aln_off + pr.n, o = out_off ? out_off[blockIdx.y] : a;
    node_mean_col(i, tensors + a * d, coords + a * 3, weights + a, tensors + b * d, coords + b * 3, weights + b, d, aln1 + a, aln2 + a,
                  xf2 + (long long)blockIdx.y * XF, t_out + o * d, c_out + o * 3, w_out + o);
}

// Sequence pool of the device-resident progressive alignment (crt_msa_*): copies the rows of each listed sequence from the pool
// into the packed chain set of a level.  tab[q] = (pool row offset, packed row offset, length); grid (chunks, sequences).
__global__ void __launch_bounds__(256) k_pool_gather(const long long *tab, const double *pt, const double *pc, const double *pw, int d,
                                                     double *t_out, double *c_out, double *w_out)
{
    const long long src = tab[blockIdx.y * 3], dst = tab[blockIdx.y * 3 + 1], len = tab[blockIdx.y * 3 + 2];
    const long long stride = (long long)gridDim.x * blockDim.x, first = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (long long e = first; e < len * d; e += stride) t_out[dst * d + e] = pt[src * d + e];
    for (long long e = first; e < len * 3; e += stride) c_out[dst * 3 + e] = pc[src * 3 + e];
    for (long long e = first; e < len; e += stride) w_out[dst + e] = pw[src + e];
}

__global__ void __launch_bounds__(256) k_fill_value(double *p, long long n, double v)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

}  // namespace crt
