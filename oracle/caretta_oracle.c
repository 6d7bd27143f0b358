/*
 * caretta_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A scalar CPU restatement of the reference's all-vs-all pair path (TurtleTools/caretta, snapshot 0.2.0,
 * mounted at /root/reference in the build container).  It is the checker the CUDA engine is compared
 * against and the "port" CPU baseline timed by bench.py.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it; the product path never does.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function below against vectors produced by
 * running the unmodified reference (numba) in the build container (oracle/gen_golden.py, fixtures under
 * tests/golden/): bit-exact for the RBF matrices, Smith-Waterman, affine DTW and common positions,
 * <= 1e-10 for everything downstream of the Kabsch SVD (the reference uses LAPACK gesdd + BLAS gemm).
 *
 * Must be compiled with -ffp-contract=off: numba/LLVM does not contract a*b+c on this path
 * (SURVEY.md fact 8), and exp() must be glibc's (the reference reaches it through llvm.exp.f64).
 *
 * Reference lines followed (paths relative to /root/reference/caretta/):
 *   crt_o_rbf_matrix            score_functions.py:6-11, 22-51          (normalized=False)
 *   crt_o_smith_waterman        dynamic_time_warping.py:225-278
 *   crt_o_smith_waterman_score  dynamic_time_warping.py:204-222
 *   crt_o_dtw_align             dynamic_time_warping.py:7-86, 89-144, 147-184
 *   crt_o_common_positions      helper.py:12-42
 *   crt_o_kabsch                superposition_functions.py:6-35, helper.py:45-53
 *   crt_o_superpose_with_subset superposition_functions.py:38-60
 *   crt_o_apply_rotran          superposition_functions.py:63-80
 *   crt_o_rmsd                  score_functions.py:14-19
 *   crt_o_tm_score              multiple_alignment.py:59-70
 *   crt_o_pair                  multiple_alignment.py:321-349 + :164-169
 *   crt_o_pairwise_all/_list    multiple_alignment.py:158-170
 *   crt_o_rmsd_cov_tm           multiple_alignment.py:1000-1055 (superpose_first=False)
 *   crt_o_neighbor_joining      neighbor_joining.py:17-157
 *   crt_o_score_matrix          multiple_alignment.py:321-349 (full matrix; progressive_align, section 8f rank 2)
 */
#include <math.h>
#include <float.h>
#include <stdint.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------------ */
/* A1: Gaussian score matrix.  acc is a strict left-to-right sum of (x-y)*(x-y); S = exp((-gamma)*acc). */
API void crt_o_rbf_matrix(const double *x, const double *y, int n, int m, int k, double gamma, double *S)
{
    const double ng = -gamma;
    for (int a = 0; a < n; ++a) {
        const double *xa = x + (size_t)a * k;
        for (int b = 0; b < m; ++b) {
            const double *yb = y + (size_t)b * k;
            double acc = 0.0;
            for (int q = 0; q < k; ++q) {
                double t = xa[q] - yb[q];
                acc = acc + t * t;
            }
            S[(size_t)a * m + b] = exp(ng * acc);
        }
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* A2 fill shared by the two SW entry points.  H is (n+1) x (m+1), row-major, caller-allocated.       */
static void sw_fill(const double *S, int n, int m, double gap, double *H)
{
    const int W = m + 1;
    for (int j = 0; j <= m; ++j) H[j] = 0.0;
    for (int i = 1; i <= n; ++i) {
        double *row = H + (size_t)i * W;
        const double *prev = row - W;
        const double *s = S + (size_t)(i - 1) * m;
        row[0] = 0.0;
        for (int j = 1; j <= m; ++j) {
            double best = 0.0;
            double dg = prev[j - 1] + s[j - 1];
            double lf = row[j - 1] - gap;
            double up = prev[j] - gap;
            if (dg > best) best = dg;
            if (lf > best) best = lf;
            if (up > best) best = up;
            row[j] = best;
        }
    }
}

/* A3: score only = max over the whole matrix. */
API double crt_o_smith_waterman_score(const double *S, int n, int m, double gap)
{
    double *H = (double *)malloc(sizeof(double) * (size_t)(n + 1) * (m + 1));
    sw_fill(S, n, m, gap, H);
    double best = 0.0; /* the matrix starts as zeros, so the max is at least 0 */
    size_t tot = (size_t)(n + 1) * (m + 1);
    for (size_t q = 0; q < tot; ++q)
        if (H[q] > best) best = H[q];
    free(H);
    return best;
}

/* A2: align.  aln1/aln2 hold at least n+m+1 entries.  Returns 0, or -1 when no cell is > 0 (the reference
 * raises there).  Output order is ascending residue index (the reference reverses before returning). */
API int crt_o_smith_waterman(const double *S, int n, int m, double gap,
                             int64_t *aln1, int64_t *aln2, int64_t *len, double *score)
{
    const int W = m + 1;
    double *H = (double *)malloc(sizeof(double) * (size_t)(n + 1) * W);
    sw_fill(S, n, m, gap, H);
    double best = 0.0;
    int bi = -1, bj = -1;
    for (int i = 1; i <= n; ++i)
        for (int j = 1; j <= m; ++j)
            if (H[(size_t)i * W + j] > best) { best = H[(size_t)i * W + j]; bi = i; bj = j; }
    *score = best;
    *len = 0;
    if (bi < 0) { free(H); return -1; }
    int i = bi, j = bj;
    int64_t k = 0;
    while (i > 0 && j > 0) {
        double h = H[(size_t)i * W + j];
        double d = H[(size_t)(i - 1) * W + (j - 1)];
        double l = H[(size_t)i * W + (j - 1)];
        double u = H[(size_t)(i - 1) * W + j];
        if (h == 0.0) break;
        else if (h == d + S[(size_t)(i - 1) * m + (j - 1)]) { --i; --j; aln1[k] = i; aln2[k] = j; ++k; }
        else if (h == l - gap) { --j; aln1[k] = -1; aln2[k] = j; ++k; }
        else if (h == u - gap) { --i; aln1[k] = i; aln2[k] = -1; ++k; }
        /* (the reference has no final else: with consistent arithmetic one branch always fires) */
        else break;
    }
    for (int64_t a = 0, b = k - 1; a < b; ++a, --b) {
        int64_t t1 = aln1[a]; aln1[a] = aln1[b]; aln1[b] = t1;
        int64_t t2 = aln2[a]; aln2[a] = aln2[b]; aln2[b] = t2;
    }
    *len = k;
    free(H);
    return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* A4: three-state affine-gap global DP with free leading gaps.  State 0 = lower (consumes i), 1 = match,
 * 2 = upper (consumes j).  argmax ties go to the lowest index.  Optionally exposes M and B (may be NULL). */
static inline int argmax2(double a, double b) { return (b > a) ? 1 : 0; }
static inline int argmax3(double a, double b, double c)
{
    int q = 0; double v = a;
    if (b > v) { v = b; q = 1; }
    if (c > v) { q = 2; }
    return q;
}

API void crt_o_dtw_align(const double *S, int n, int m, double open, double ext,
                         int64_t *aln1, int64_t *aln2, int64_t *len, double *score,
                         double *M_out, int64_t *B_out)
{
    const double MINF = -DBL_MAX; /* np.finfo(float64).min */
    const size_t W = (size_t)(m + 1) * 3;
    double *M = (double *)calloc((size_t)(n + 1) * W, sizeof(double));
    uint8_t *B = (uint8_t *)calloc((size_t)(n + 1) * W, 1);
#define MM(i, j, s) M[(size_t)(i) * W + (size_t)(j) * 3 + (s)]
#define BB(i, j, s) B[(size_t)(i) * W + (size_t)(j) * 3 + (s)]
    for (int i = 0; i <= n; ++i) for (int s = 0; s < 3; ++s) MM(i, 0, s) = MINF;
    for (int j = 0; j <= m; ++j) for (int s = 0; s < 3; ++s) MM(0, j, s) = MINF;
    for (int s = 0; s < 3; ++s) MM(0, 0, s) = 0.0;
    for (int i = 1; i <= n; ++i) {
        MM(i, 0, 0) = 0.0; MM(i, 0, 1) = 0.0; MM(i, 0, 2) = MINF - open;
        BB(i, 0, 0) = BB(i, 0, 1) = BB(i, 0, 2) = 0;
    }
    for (int j = 1; j <= m; ++j) {
        MM(0, j, 0) = MINF - open; MM(0, j, 1) = 0.0; MM(0, j, 2) = 0.0;
        BB(0, j, 0) = BB(0, j, 1) = BB(0, j, 2) = 1;
    }
    for (int i = 1; i <= n; ++i) {
        for (int j = 1; j <= m; ++j) {
            double l0 = MM(i - 1, j, 0) - ext, l1 = MM(i - 1, j, 1) - open;
            int ql = argmax2(l0, l1);
            double lower = ql ? l1 : l0;
            MM(i, j, 0) = lower; BB(i, j, 0) = (uint8_t)ql;
            double u0 = MM(i, j - 1, 1) - open, u1 = MM(i, j - 1, 2) - ext;
            int qu = argmax2(u0, u1);
            double upper = qu ? u1 : u0;
            MM(i, j, 2) = upper; BB(i, j, 2) = (uint8_t)(qu + 1);
            double dg = MM(i - 1, j - 1, 1) + S[(size_t)(i - 1) * m + (j - 1)];
            int q = argmax3(lower, dg, upper);
            MM(i, j, 1) = (q == 0) ? lower : (q == 1 ? dg : upper);
            BB(i, j, 1) = (uint8_t)q;
        }
    }
    int dir = argmax3(MM(n, m, 0), MM(n, m, 1), MM(n, m, 2));
    *score = MM(n, m, dir);
    int64_t k = 0;
    int a = n, b = m;
    while (!(a == 0 && b == 0)) {
        if (b == 0) { --a; aln1[k] = a; aln2[k] = -1; ++k; }
        else if (a == 0) { --b; aln1[k] = -1; aln2[k] = b; ++k; }
        else if (dir == 0) { dir = BB(a, b, 0); --a; aln1[k] = a; aln2[k] = -1; ++k; }
        else if (dir == 1) {
            dir = BB(a, b, 1);
            if (dir == 1) { --a; --b; aln1[k] = a; aln2[k] = b; ++k; }
        } else { dir = BB(a, b, 2); --b; aln1[k] = -1; aln2[k] = b; ++k; }
    }
    for (int64_t x = 0, y = k - 1; x < y; ++x, --y) {
        int64_t t1 = aln1[x]; aln1[x] = aln1[y]; aln1[y] = t1;
        int64_t t2 = aln2[x]; aln2[x] = aln2[y]; aln2[y] = t2;
    }
    *len = k;
    if (M_out) memcpy(M_out, M, sizeof(double) * (size_t)(n + 1) * W);
    if (B_out) for (size_t q = 0; q < (size_t)(n + 1) * W; ++q) B_out[q] = B[q];
#undef MM
#undef BB
    free(M); free(B);
}

/* ------------------------------------------------------------------------------------------------ */
/* A5 */
API int64_t crt_o_common_positions(const int64_t *aln1, const int64_t *aln2, int64_t len,
                                   int64_t *pos1, int64_t *pos2)
{
    int64_t c = 0;
    for (int64_t q = 0; q < len; ++q)
        if (aln1[q] != -1 && aln2[q] != -1) { pos1[c] = aln1[q]; pos2[c] = aln2[q]; ++c; }
    return c;
}

/* ------------------------------------------------------------------------------------------------ */
/* 3x3 SVD by one-sided Jacobi (the reference calls LAPACK gesdd; results agree to rounding).
 * A = U diag(s) Vt, s sorted descending, like numpy.linalg.svd. */
static void svd3(const double A[9], double U[9], double s[3], double Vt[9])
{
    double W[9], V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    memcpy(W, A, sizeof(W));
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                double alpha = 0, beta = 0, gam = 0;
                for (int r = 0; r < 3; ++r) {
                    alpha += W[r * 3 + p] * W[r * 3 + p];
                    beta += W[r * 3 + q] * W[r * 3 + q];
                    gam += W[r * 3 + p] * W[r * 3 + q];
                }
                if (gam == 0.0) continue;
                double lim = sqrt(alpha * beta);
                if (fabs(gam) <= 1e-300 || fabs(gam) <= 2.2e-16 * lim) continue;
                off = fmax(off, fabs(gam) / (lim > 0 ? lim : 1));
                double zeta = (beta - alpha) / (2.0 * gam);
                double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
                for (int r = 0; r < 3; ++r) {
                    double wp = W[r * 3 + p], wq = W[r * 3 + q];
                    W[r * 3 + p] = c * wp - sn * wq; W[r * 3 + q] = sn * wp + c * wq;
                    double vp = V[r * 3 + p], vq = V[r * 3 + q];
                    V[r * 3 + p] = c * vp - sn * vq; V[r * 3 + q] = sn * vp + c * vq;
                }
            }
        if (off == 0.0) break;
    }
    double nrm[3];
    int ord[3] = {0, 1, 2};
    for (int q = 0; q < 3; ++q)
        nrm[q] = sqrt(W[q] * W[q] + W[3 + q] * W[3 + q] + W[6 + q] * W[6 + q]);
    for (int a = 0; a < 2; ++a)
        for (int b = a + 1; b < 3; ++b)
            if (nrm[ord[b]] > nrm[ord[a]]) { int t = ord[a]; ord[a] = ord[b]; ord[b] = t; }
    double Uc[3][3];
    for (int q = 0; q < 3; ++q) {
        int o = ord[q];
        s[q] = nrm[o];
        for (int r = 0; r < 3; ++r) {
            Vt[q * 3 + r] = V[r * 3 + o];
            Uc[q][r] = (nrm[o] > 0) ? W[r * 3 + o] / nrm[o] : 0.0;
        }
    }
    /* rank-deficient input: complete U to an orthonormal basis (any completion gives the same Kabsch R
     * after the reflection fix as long as s[1] > s[2]). */
    if (!(s[2] > 1e-14 * (s[0] > 0 ? s[0] : 1))) {
        if (!(s[1] > 1e-14 * (s[0] > 0 ? s[0] : 1))) {
            if (!(s[0] > 0)) { Uc[0][0] = 1; Uc[0][1] = 0; Uc[0][2] = 0; }
            double ax[3] = {0, 0, 0};
            int mn = 0;
            for (int r = 1; r < 3; ++r) if (fabs(Uc[0][r]) < fabs(Uc[0][mn])) mn = r;
            ax[mn] = 1.0;
            double dp = ax[0] * Uc[0][0] + ax[1] * Uc[0][1] + ax[2] * Uc[0][2];
            double nn = 0;
            for (int r = 0; r < 3; ++r) { Uc[1][r] = ax[r] - dp * Uc[0][r]; nn += Uc[1][r] * Uc[1][r]; }
            nn = sqrt(nn);
            for (int r = 0; r < 3; ++r) Uc[1][r] /= nn;
        }
        Uc[2][0] = Uc[0][1] * Uc[1][2] - Uc[0][2] * Uc[1][1];
        Uc[2][1] = Uc[0][2] * Uc[1][0] - Uc[0][0] * Uc[1][2];
        Uc[2][2] = Uc[0][0] * Uc[1][1] - Uc[0][1] * Uc[1][0];
    }
    for (int q = 0; q < 3; ++q)
        for (int r = 0; r < 3; ++r) U[r * 3 + q] = Uc[q][r];
}

static double det3(const double A[9])
{
    return A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6])
         + A[2] * (A[3] * A[7] - A[4] * A[6]);
}

static void mean_axis0(const double *x, int64_t c, double mu[3])
{
    for (int a = 0; a < 3; ++a) {
        double acc = 0.0;
        for (int64_t r = 0; r < c; ++r) acc += x[r * 3 + a];
        mu[a] = acc / (double)c;
    }
}

/* A6: Kabsch.  Row-vector convention: x2 @ R + t ~= x1.  R is row-major 3x3. */
API void crt_o_kabsch(const double *x1, const double *x2, int64_t c, double *R, double *t)
{
    double m1[3], m2[3], C[9] = {0};
    mean_axis0(x1, c, m1);
    mean_axis0(x2, c, m2);
    for (int64_t r = 0; r < c; ++r)
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b)
                C[a * 3 + b] += (x2[r * 3 + a] - m2[a]) * (x1[r * 3 + b] - m1[b]);
    double U[9], s[3], Vt[9];
    svd3(C, U, s, Vt);
    if (det3(U) * det3(Vt) < 0)
        for (int r = 0; r < 3; ++r) U[r * 3 + 2] = -U[r * 3 + 2];
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
            double acc = 0.0;
            for (int q = 0; q < 3; ++q) acc += U[a * 3 + q] * Vt[q * 3 + b];
            R[a * 3 + b] = acc;
        }
    for (int b = 0; b < 3; ++b)
        t[b] = m1[b] - (m2[0] * R[0 * 3 + b] + m2[1] * R[1 * 3 + b] + m2[2] * R[2 * 3 + b]);
}

API void crt_o_apply_rotran(const double *x, int64_t n, const double *R, const double *t, double *out)
{
    for (int64_t r = 0; r < n; ++r)
        for (int b = 0; b < 3; ++b)
            out[r * 3 + b] = (x[r * 3 + 0] * R[0 * 3 + b] + x[r * 3 + 1] * R[1 * 3 + b]
                              + x[r * 3 + 2] * R[2 * 3 + b]) + t[b];
}

/* superposition_functions.py:38-60: both chains end up in the frame of the common centroids. */
API void crt_o_superpose_with_subset(const double *c1, int64_t n, const double *c2, int64_t m,
                                     const double *k1, const double *k2, int64_t c,
                                     double *out1, double *out2, double *R_out)
{
    double R[9], t[3], m1[3], m2[3];
    crt_o_kabsch(k1, k2, c, R, t);
    mean_axis0(k1, c, m1);
    mean_axis0(k2, c, m2);
    for (int64_t r = 0; r < n; ++r)
        for (int a = 0; a < 3; ++a) out1[r * 3 + a] = c1[r * 3 + a] - m1[a];
    for (int64_t r = 0; r < m; ++r) {
        double y0 = c2[r * 3 + 0] - m2[0], y1 = c2[r * 3 + 1] - m2[1], y2 = c2[r * 3 + 2] - m2[2];
        for (int b = 0; b < 3; ++b)
            out2[r * 3 + b] = y0 * R[0 * 3 + b] + y1 * R[1 * 3 + b] + y2 * R[2 * 3 + b];
    }
    if (R_out) memcpy(R_out, R, sizeof(R));
}

/* A7 */
API double crt_o_rmsd(const double *x, const double *y, int64_t c)
{
    double acc = 0.0;
    for (int64_t q = 0; q < c * 3; ++q) { double t = x[q] - y[q]; acc += t * t; }
    return sqrt(acc / (double)c);
}

/* multiple_alignment.py:59-70, quirks kept: d = 1.24*(l-15)/3 - 1.8 and a *signed coordinate sum*. */
API double crt_o_tm_score(const double *x, const double *y, int64_t c, int64_t l1, int64_t l2)
{
    double d1 = 1.24 * (double)(l1 - 15) / 3 - 1.8;
    double d2 = 1.24 * (double)(l2 - 15) / 3 - 1.8;
    double s1 = 0.0, s2 = 0.0;
    for (int64_t r = 0; r < c; ++r) {
        double sm = 0.0;
        for (int a = 0; a < 3; ++a) sm += x[r * 3 + a] - y[r * 3 + a];
        double q1 = sm / d1, q2 = sm / d2;
        s1 += 1 / (1 + q1 * q1);
        s2 += 1 / (1 + q2 * q2);
    }
    double t1 = (1.0 / (double)l1) * s1, t2 = (1.0 / (double)l2) * s2;
    return t1 > t2 ? t1 : t2;
}

/* ------------------------------------------------------------------------------------------------ */
/* One pair of the hot loop (multiple_alignment.py:321-349 then :164-169).
 * Optional outputs (NULL to skip): stage-1 paths (capacity n+m+1), their length, number of common
 * positions, Kabsch rotation, RMSD / TM over the stage-1 matched residues (engine by-products, validated
 * against the same composition of reference functions), stage-1 SW score.
 * status bits: 1 = <= 3 common positions (superposition skipped), 2 = stage-1 matrix had no cell > 0. */
API int crt_o_pair(const double *t1, const double *c1, int n, const double *t2, const double *c2, int m,
                   int d, double gamma_t, double gamma_c,
                   double *score, int64_t *aln1, int64_t *aln2, int64_t *aln_len, int64_t *ncommon,
                   double *R_out, double *rmsd, double *tm, double *score1)
{
    int status = 0;
    double *S = (double *)malloc(sizeof(double) * (size_t)n * m);
    int64_t *a1 = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n + m + 1) * 4);
    int64_t *a2 = a1 + (n + m + 1), *p1 = a2 + (n + m + 1), *p2 = p1 + (n + m + 1);
    double *w1 = (double *)malloc(sizeof(double) * (size_t)(n + m) * 3 * 2);
    double *w2 = w1 + (size_t)n * 3, *k1 = w2 + (size_t)m * 3, *k2 = k1 + (size_t)(n < m ? n : m) * 3;
    int64_t len = 0, c = 0;
    double sc1 = 0.0;
    crt_o_rbf_matrix(t1, t2, n, m, d, gamma_t, S);
    if (crt_o_smith_waterman(S, n, m, 0.0, a1, a2, &len, &sc1) != 0) status |= 2;
    c = crt_o_common_positions(a1, a2, len, p1, p2);
    for (int64_t q = 0; q < c; ++q)
        for (int a = 0; a < 3; ++a) { k1[q * 3 + a] = c1[p1[q] * 3 + a]; k2[q * 3 + a] = c2[p2[q] * 3 + a]; }
    double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    if (c <= 3) {
        status |= 1;
        memcpy(w1, c1, sizeof(double) * (size_t)n * 3);
        memcpy(w2, c2, sizeof(double) * (size_t)m * 3);
    } else {
        crt_o_superpose_with_subset(c1, n, c2, m, k1, k2, c, w1, w2, R);
    }
    crt_o_rbf_matrix(w1, w2, n, m, 3, gamma_c, S);
    *score = crt_o_smith_waterman_score(S, n, m, 0.0);
    if (rmsd || tm) {
        double rr = 0.0, tt = 0.0;
        if (c >= 1) {
            double Rk[9], tk[3];
            double *k2r = (double *)malloc(sizeof(double) * (size_t)c * 3);
            if (c > 3) { crt_o_kabsch(k1, k2, c, Rk, tk); crt_o_apply_rotran(k2, c, Rk, tk, k2r); }
            else memcpy(k2r, k2, sizeof(double) * (size_t)c * 3);
            rr = crt_o_rmsd(k1, k2r, c);
            tt = crt_o_tm_score(k1, k2r, c, n, m);
            free(k2r);
        }
        if (rmsd) *rmsd = rr;
        if (tm) *tm = tt;
    }
    if (aln1) memcpy(aln1, a1, sizeof(int64_t) * (size_t)len);
    if (aln2) memcpy(aln2, a2, sizeof(int64_t) * (size_t)len);
    if (aln_len) *aln_len = len;
    if (ncommon) *ncommon = c;
    if (R_out) memcpy(R_out, R, sizeof(R));
    if (score1) *score1 = sc1;
    free(S); free(a1); free(w1);
    return status;
}

/* Protein.score_function(flexible=False), multiple_alignment.py:321-349: the full n x m coordinate score matrix of a pair
 * (what progressive_align feeds to dtw_align, :203-214).  Returns the status bits of crt_o_pair. */
API int crt_o_score_matrix(const double *t1, const double *c1, int n, const double *t2, const double *c2, int m,
                           int d, double gamma_t, double gamma_c, double *S_out)
{
    int status = 0;
    int64_t *a1 = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n + m + 1) * 4);
    int64_t *a2 = a1 + (n + m + 1), *p1 = a2 + (n + m + 1), *p2 = p1 + (n + m + 1);
    double *w1 = (double *)malloc(sizeof(double) * (size_t)(n + m) * 3 * 2);
    double *w2 = w1 + (size_t)n * 3, *k1 = w2 + (size_t)m * 3, *k2 = k1 + (size_t)(n < m ? n : m) * 3;
    int64_t len = 0, c = 0;
    double sc1 = 0.0;
    crt_o_rbf_matrix(t1, t2, n, m, d, gamma_t, S_out);
    if (crt_o_smith_waterman(S_out, n, m, 0.0, a1, a2, &len, &sc1) != 0) status |= 2;
    c = crt_o_common_positions(a1, a2, len, p1, p2);
    for (int64_t q = 0; q < c; ++q)
        for (int a = 0; a < 3; ++a) { k1[q * 3 + a] = c1[p1[q] * 3 + a]; k2[q * 3 + a] = c2[p2[q] * 3 + a]; }
    if (c <= 3) {
        status |= 1;
        memcpy(w1, c1, sizeof(double) * (size_t)n * 3);
        memcpy(w2, c2, sizeof(double) * (size_t)m * 3);
    } else {
        crt_o_superpose_with_subset(c1, n, c2, m, k1, k2, c, w1, w2, NULL);
    }
    crt_o_rbf_matrix(w1, w2, n, m, 3, gamma_c, S_out);
    free(a1); free(w1);
    return status;
}

/* Explicit pair list over packed chains (coords [sumL,3], tensors [sumL,d], offsets [N+1]).
 * Per-pair outputs; any of rmsd/tm/ncommon/status may be NULL.  nthreads <= 0 -> all cores. */
API void crt_o_pairwise_list(const double *coords, const double *tensors, const int64_t *offsets, int d,
                             const int32_t *pi, const int32_t *pj, int64_t n_pairs,
                             double gamma_t, double gamma_c, int nthreads,
                             double *score, double *rmsd, double *tm, int32_t *ncommon, int32_t *status)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t q = 0; q < n_pairs; ++q) {
        int i = pi[q], j = pj[q];
        int n = (int)(offsets[i + 1] - offsets[i]), m = (int)(offsets[j + 1] - offsets[j]);
        double sc, rr, tt; int64_t c;
        int st = crt_o_pair(tensors + offsets[i] * d, coords + offsets[i] * 3, n,
                            tensors + offsets[j] * d, coords + offsets[j] * 3, m, d, gamma_t, gamma_c,
                            &sc, NULL, NULL, NULL, &c, NULL, (rmsd || tm) ? &rr : NULL,
                            (rmsd || tm) ? &tt : NULL, NULL);
        score[q] = sc;
        if (rmsd) rmsd[q] = rr;
        if (tm) tm[q] = tt;
        if (ncommon) ncommon[q] = (int32_t)c;
        if (status) status[q] = st;
    }
}

/* The driver loop of multiple_alignment.py:158-170: dense symmetric [N,N], diagonal left at 0. */
API void crt_o_pairwise_all(const double *coords, const double *tensors, const int64_t *offsets, int N, int d,
                            double gamma_t, double gamma_c, int nthreads, double *out /* [N*N] */)
{
    int64_t np = (int64_t)N * (N - 1) / 2;
    int32_t *pi = (int32_t *)malloc(sizeof(int32_t) * (size_t)(np > 0 ? np : 1) * 2), *pj = pi + np;
    double *sc = (double *)malloc(sizeof(double) * (size_t)(np > 0 ? np : 1));
    int64_t q = 0;
    for (int i = 0; i < N - 1; ++i) for (int j = i + 1; j < N; ++j) { pi[q] = i; pj[q] = j; ++q; }
    crt_o_pairwise_list(coords, tensors, offsets, d, pi, pj, np, gamma_t, gamma_c, nthreads,
                        sc, NULL, NULL, NULL, NULL);
    memset(out, 0, sizeof(double) * (size_t)N * N);
    for (q = 0; q < np; ++q) { out[(size_t)pi[q] * N + pj[q]] = sc[q]; out[(size_t)pj[q] * N + pi[q]] = sc[q]; }
    free(pi); free(sc);
}

/* multiple_alignment.py:1000-1055 with superpose_first=False.  aln is [N, A] int64 (-1 = gap).
 * Returns the number of pairs with < 3 common positions (the reference asserts there); those entries are
 * left at the diagonal defaults. */
API int crt_o_rmsd_cov_tm(const int64_t *aln, int N, int64_t A, const double *coords, const int64_t *offsets,
                          double *rmsd, double *cov, double *tm)
{
    int bad = 0;
    for (int64_t q = 0; q < (int64_t)N * N; ++q) { rmsd[q] = 0.0; cov[q] = 1.0; tm[q] = 1.0; }
    int64_t *p1 = (int64_t *)malloc(sizeof(int64_t) * (size_t)(A + 1) * 2), *p2 = p1 + (A + 1);
    double *k1 = (double *)malloc(sizeof(double) * (size_t)(A + 1) * 9), *k2 = k1 + (A + 1) * 3,
           *k2r = k2 + (A + 1) * 3;
    for (int i = 0; i < N - 1; ++i)
        for (int j = i + 1; j < N; ++j) {
            int64_t c = crt_o_common_positions(aln + (size_t)i * A, aln + (size_t)j * A, A, p1, p2);
            if (c < 3) { ++bad; continue; }
            const double *ci = coords + offsets[i] * 3, *cj = coords + offsets[j] * 3;
            for (int64_t r = 0; r < c; ++r)
                for (int a = 0; a < 3; ++a) { k1[r * 3 + a] = ci[p1[r] * 3 + a]; k2[r * 3 + a] = cj[p2[r] * 3 + a]; }
            double R[9], t[3];
            crt_o_kabsch(k1, k2, c, R, t);
            crt_o_apply_rotran(k2, c, R, t, k2r);
            double rr = crt_o_rmsd(k1, k2r, c);
            double cv = (double)c / (double)A;
            double tt = crt_o_tm_score(k1, k2r, c, offsets[i + 1] - offsets[i], offsets[j + 1] - offsets[j]);
            rmsd[(size_t)i * N + j] = rmsd[(size_t)j * N + i] = rr;
            cov[(size_t)i * N + j] = cov[(size_t)j * N + i] = cv;
            tm[(size_t)i * N + j] = tm[(size_t)j * N + i] = tt;
        }
    free(p1); free(k1);
    return bad;
}

/* ------------------------------------------------------------------------------------------------ */
/* neighbor_joining.py:17-157.  The reference recomputes np.sum(distance_matrix[i, :]) inside the O(n^2) scan of
 * _find_join_nodes (:118-125); numba's np.sum is a plain left-to-right loop, so caching the n row sums per iteration
 * gives the identical values.  Q is evaluated in the reference's order ((n-2)*d - sum_i) - sum_j, the minimum is the
 * first strict minimum in row-major order starting from +inf (:127-129), the matrix is rebuilt with the new node at
 * index 0 and the remaining nodes in their previous order (:59-77).  tree: [2N-3][2] uint64, bl: [2N-3] float64.
 * Returns the number of rows written, or -1 if N < 3 (the reference indexes out of range there). */
API int64_t crt_o_neighbor_joining(const double *D0, int N, uint64_t *tree, double *bl)
{
    if (N < 3) return -1;
    int n = N;
    double *D = (double *)malloc(sizeof(double) * (size_t)N * N), *Dn = (double *)malloc(sizeof(double) * (size_t)N * N);
    double *S = (double *)malloc(sizeof(double) * (size_t)N);
    int64_t *ti = (int64_t *)malloc(sizeof(int64_t) * (size_t)N), *tn = (int64_t *)malloc(sizeof(int64_t) * (size_t)N);
    int *idx = (int *)malloc(sizeof(int) * (size_t)N);
    memcpy(D, D0, sizeof(double) * (size_t)N * N);
    for (int i = 0; i < N; ++i) ti[i] = i;
    int64_t index = 0, n_inter = 0;
    while (n > 3) {
        for (int i = 0; i < n; ++i) { double acc = 0.0; for (int j = 0; j < n; ++j) acc = acc + D[(size_t)i * n + j]; S[i] = acc; }
        int mi = 0, mj = 0;
        double min_q = INFINITY;
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j)
                if (i != j) {
                    const double q = ((double)(n - 2) * D[(size_t)i * n + j] - S[i]) - S[j];
                    if (q < min_q) { mi = i; mj = j; min_q = q; }
                }
        const double dij = D[(size_t)mi * n + mj];
        const double di = 0.5 * dij + (0.5 / (double)(n - 2)) * (S[mi] - S[mj]);      /* _find_branch_length :137-157 */
        const double dj = dij - di;
        const int64_t node = n_inter + N;
        ++n_inter;
        tree[2 * index] = (uint64_t)ti[mi]; tree[2 * index + 1] = (uint64_t)node; bl[index] = di; ++index;
        tree[2 * index] = (uint64_t)ti[mj]; tree[2 * index + 1] = (uint64_t)node; bl[index] = dj; ++index;
        int cnt = 0;
        for (int i = 0; i < n; ++i) if (i != mi && i != mj) idx[cnt++] = i;
        const int nn = n - 1;
        Dn[0] = 0.0;
        for (int a = 0; a < cnt; ++a)
            for (int b = 0; b < cnt; ++b) Dn[(size_t)(a + 1) * nn + (b + 1)] = D[(size_t)idx[a] * n + idx[b]];
        for (int a = 0; a < cnt; ++a) {
            const double v = 0.5 * ((D[(size_t)mi * n + idx[a]] + D[(size_t)mj * n + idx[a]]) - dij);
            Dn[a + 1] = v; Dn[(size_t)(a + 1) * nn] = v;
        }
        tn[0] = node;
        for (int a = 0; a < cnt; ++a) tn[a + 1] = ti[idx[a]];
        { double *t = D; D = Dn; Dn = t; }
        { int64_t *t = ti; ti = tn; tn = t; }
        n = nn;
    }
    {   /* last three nodes, :80-98 */
        double s1 = 0.0, s2 = 0.0;
        for (int j = 0; j < n; ++j) { s1 = s1 + D[(size_t)1 * n + j]; s2 = s2 + D[(size_t)2 * n + j]; }
        const double d12 = D[(size_t)1 * n + 2];
        const double di = 0.5 * d12 + (0.5 / (double)(n - 2)) * (s1 - s2);
        const double dj = d12 - di;
        const int64_t node = n_inter + N;
        tree[2 * index] = (uint64_t)ti[1]; tree[2 * index + 1] = (uint64_t)node; bl[index] = di; ++index;
        tree[2 * index] = (uint64_t)ti[2]; tree[2 * index + 1] = (uint64_t)node; bl[index] = dj; ++index;
        tree[2 * index] = (uint64_t)ti[0]; tree[2 * index + 1] = (uint64_t)node;
        bl[index] = 0.5 * ((D[(size_t)1 * n + 0] + D[(size_t)2 * n + 0]) - D[(size_t)1 * n + 2]); ++index;
    }
    free(D); free(Dn); free(S); free(ti); free(tn); free(idx);
    return index;
}

/* ------------------------------------------------------------------------------------------------ */
/* fp32 models of the production arithmetic, used only to study how far an fp32 DP can agree with the fp64
 * reference (tests/test_fp32_model.py).  Not a restatement of reference code.
 *   variant 0: direct-difference RBF, absolute-value DP   H = max(Hd + S, Hl, Hu)
 *   variant 1: dot-product RBF (|a|^2 + |b|^2 - 2ab folded into an FMA chain), difference-form DP
 *              d = max(S, a, b), u = d - a, v = d - b   with a = vertical diff of the left cell, b = horizontal
 *              diff of the upper cell; all quantities stay in [0, 1] so fp32 keeps ~100x more absolute resolution.
 * Traceback codes are taken at fill time exactly as the CUDA kernels do (diag if d==S, else left if d==a, else up).
 */
static float g_f32_flush = 0.f;   /* model of a flush-to-zero threshold on S (study only) */
API void crt_o_set_f32_flush(double v) { g_f32_flush = (float)v; }

API int crt_o_pair_f32model(const double *t1, const double *c1, int n, const double *t2, const double *c2,
                            int m, int d, double gamma_t, double gamma_c, int variant,
                            double *score, int64_t *aln1, int64_t *aln2, int64_t *aln_len)
{
    (void)c1; (void)c2; (void)gamma_c;
    float *S = (float *)malloc(sizeof(float) * (size_t)n * m);
    uint8_t *code = (uint8_t *)calloc((size_t)(n + 1) * (m + 1), 1);
    const int W = m + 1;
    const float g2 = (float)(gamma_t * 1.4426950408889634);
    if (variant == 0) {
        for (int a = 0; a < n; ++a)
            for (int b = 0; b < m; ++b) {
                float acc = 0.f;
                for (int q = 0; q < d; ++q) {
                    float t = (float)t1[(size_t)a * d + q] - (float)t2[(size_t)b * d + q];
                    acc = fmaf(t, t, acc);
                }
                S[(size_t)a * m + b] = exp2f(-g2 * acc);
            }
    } else {
        float *ap = (float *)malloc(sizeof(float) * (size_t)(n + m) * (d + 1));
        float *bp = ap + (size_t)n * (d + 1);
        for (int a = 0; a < n; ++a) {
            double nn = 0;
            for (int q = 0; q < d; ++q) { double v = t1[(size_t)a * d + q]; nn += v * v; ap[(size_t)a * (d + 1) + q] = (float)(2.0 * gamma_t * 1.4426950408889634 * v); }
            ap[(size_t)a * (d + 1) + d] = (float)(-gamma_t * 1.4426950408889634 * nn);
        }
        for (int b = 0; b < m; ++b) {
            double nn = 0;
            for (int q = 0; q < d; ++q) { double v = t2[(size_t)b * d + q]; nn += v * v; bp[(size_t)b * (d + 1) + q] = (float)v; }
            bp[(size_t)b * (d + 1) + d] = (float)(-gamma_t * 1.4426950408889634 * nn);
        }
        for (int a = 0; a < n; ++a)
            for (int b = 0; b < m; ++b) {
                float acc = ap[(size_t)a * (d + 1) + d] + bp[(size_t)b * (d + 1) + d];
                for (int q = 0; q < d; ++q) acc = fmaf(ap[(size_t)a * (d + 1) + q], bp[(size_t)b * (d + 1) + q], acc);
                S[(size_t)a * m + b] = exp2f(acc);
                if (S[(size_t)a * m + b] < g_f32_flush) S[(size_t)a * m + b] = 0.f;
            }
        free(ap);
    }
    int bi = -1, bj = -1;
    float best = 0.f;
    if (variant == 0) {
        float *H = (float *)calloc((size_t)(n + 1) * W, sizeof(float));
        for (int i = 1; i <= n; ++i)
            for (int j = 1; j <= m; ++j) {
                float dg = H[(size_t)(i - 1) * W + j - 1] + S[(size_t)(i - 1) * m + j - 1];
                float lf = H[(size_t)i * W + j - 1], up = H[(size_t)(i - 1) * W + j];
                float h = dg > lf ? dg : lf; h = h > up ? h : up;
                H[(size_t)i * W + j] = h;
                code[(size_t)i * W + j] = (h == 0.f) ? 0 : (h == dg ? 1 : (h == lf ? 2 : 3));
            }
        for (int i = 1; i <= n; ++i)
            for (int j = 1; j <= m; ++j)
                if (H[(size_t)i * W + j] > best) { best = H[(size_t)i * W + j]; bi = i; bj = j; }
        free(H);
    } else {
        float *u = (float *)calloc((size_t)W, sizeof(float));     /* horizontal diffs of the previous row */
        uint8_t *lefteq = (uint8_t *)calloc((size_t)(n + 1) * W, 1);
        double hsum = 0.0;
        int istar = -1;
        for (int i = 1; i <= n; ++i) {
            float a = 0.f;                                          /* v[i][0] = 0 */
            for (int j = 1; j <= m; ++j) {
                float s = S[(size_t)(i - 1) * m + j - 1], b = u[j];
                float dd = s > a ? s : a; dd = dd > b ? dd : b;
                code[(size_t)i * W + j] = (dd == s) ? 1 : (dd == a ? 2 : 3);
                lefteq[(size_t)i * W + j] = (dd == a);
                u[j] = dd - a;
                a = dd - b;
            }
            if (a > 0.f) istar = i;                                 /* H[i][m] > H[i-1][m] */
            hsum += a;
        }
        best = (float)hsum;
        if (istar > 0) { bi = istar; bj = m; while (bj > 1 && lefteq[(size_t)bi * W + bj]) --bj; }
        free(u); free(lefteq);
    }
    *score = best; *aln_len = 0;
    if (bi < 0) { free(S); free(code); return -1; }
    int i = bi, j = bj; int64_t k = 0;
    while (i > 0 && j > 0) {
        int cd = code[(size_t)i * W + j];
        if (cd == 0) break;
        else if (cd == 1) { --i; --j; aln1[k] = i; aln2[k] = j; ++k; }
        else if (cd == 2) { --j; aln1[k] = -1; aln2[k] = j; ++k; }
        else { --i; aln1[k] = i; aln2[k] = -1; ++k; }
    }
    for (int64_t a = 0, b = k - 1; a < b; ++a, --b) {
        int64_t x = aln1[a]; aln1[a] = aln1[b]; aln1[b] = x;
        x = aln2[a]; aln2[a] = aln2[b]; aln2[b] = x;
    }
    *aln_len = k;
    free(S); free(code);
    return 0;
}

API int crt_o_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------------ */
/* Consumers of the multiple alignment.                               */
/* ------------------------------------------------------------------------------------------------ */

/* make_coverage_gap_distance_matrix, multiple_alignment.py:45-56: for every i the columns where i has a residue;
 * num_gaps(i, j) = how many of them are gaps in j; distance = num_gaps / length_i; aligning = length_i - num_gaps.
 * Returns -1 when some protein has no residue at all (the reference divides by zero there). */
API int crt_o_coverage_gap_matrix(const int64_t *aln, int N, int64_t A, double *distance, int32_t *aligning)
{
    int64_t *idx = (int64_t *)malloc(sizeof(int64_t) * (size_t)(A > 0 ? A : 1));
    int rc = 0;
    for (int i = 0; i < N && rc == 0; ++i) {
        int64_t len = 0;
        for (int64_t q = 0; q < A; ++q)
            if (aln[(size_t)i * A + q] != -1) idx[len++] = q;
        if (len == 0) { rc = -1; break; }
        for (int j = 0; j < N; ++j) {
            int64_t gaps = 0;
            for (int64_t k = 0; k < len; ++k) gaps += aln[(size_t)j * A + idx[k]] == -1;
            distance[(size_t)i * N + j] = (double)gaps / (double)len;
            aligning[(size_t)i * N + j] = (int32_t)(len - gaps);
        }
    }
    free(idx);
    return rc;
}

/* helper.write_distance_matrix, helper.py:183-203: f"{len(names)}\n" then f"{name} {' '.join(f'{x:.4f}' ...)}\n".
 * Python's '.4f' is the correctly rounded decimal expansion, like glibc's printf("%.4f"); the one difference is that
 * Python prints "nan" for every NaN where printf prints "-nan" when the sign bit is set.
 * out = NULL: only returns the length.  names: packed bytes, name_off [n_rows + 1]. */
API int64_t crt_o_format_matrix(const double *M, int n_rows, int n_cols, const char *names, const int64_t *name_off,
                                char *out)
{
    char buf[400];
    int64_t pos = 0;
    int n = snprintf(buf, sizeof(buf), "%d\n", n_rows);
    if (out) memcpy(out + pos, buf, (size_t)n);
    pos += n;
    for (int i = 0; i < n_rows; ++i) {
        int64_t nl = name_off[i + 1] - name_off[i];
        if (out) { memcpy(out + pos, names + name_off[i], (size_t)nl); out[pos + nl] = ' '; }
        pos += nl + 1;
        for (int j = 0; j < n_cols; ++j) {
            double x = M[(size_t)i * n_cols + j];
            if (x != x) n = snprintf(buf, sizeof(buf), "nan");
            else n = snprintf(buf, sizeof(buf), "%.4f", x);
            if (j) { if (out) out[pos] = ' '; ++pos; }
            if (out) memcpy(out + pos, buf, (size_t)n);
            pos += n;
        }
        if (out) out[pos] = '\n';
        ++pos;
    }
    return pos;
}

/* make_count_matrix, multiple_alignment.py:128-134.  Returns -1 on an index outside [0, K). */
API int crt_o_count_matrix(const int64_t *idx, const int64_t *off, int N, int K, double *out)
{
    memset(out, 0, sizeof(double) * (size_t)N * K);
    for (int i = 0; i < N; ++i)
        for (int64_t q = off[i]; q < off[i + 1]; ++q) {
            if (idx[q] < 0 || idx[q] >= K) return -1;
            out[(size_t)i * K + idx[q]] += 1.0;
        }
    return 0;
}

/* braycurtis, multiple_alignment.py:137-145: np.abs(a - b).sum() / np.abs(a + b).sum(), sums in index order. */
API void crt_o_braycurtis(const double *A, int n1, const double *B, int n2, int K, double *out)
{
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n1; ++i)
        for (int j = 0; j < n2; ++j) {
            double num = 0.0, den = 0.0;
            for (int k = 0; k < K; ++k) {
                num += fabs(A[(size_t)i * K + k] - B[(size_t)j * K + k]);
                den += fabs(A[(size_t)i * K + k] + B[(size_t)j * K + k]);
            }
            out[(size_t)i * n2 + j] = num / den;
        }
}
