/* tie_study.c -- STUDY TOOL (not product, not oracle): where does the fp32 difference-form stage-1 DP take another path
 * than the reference's float64 Smith-Waterman, and which cheap per-cell criterion flags those pairs?
 *
 * For a sample of pairs of a packed chain set it runs (a) the reference arithmetic in float64 (score_functions.py:6-11,
 * dynamic_time_warping.py:225-278: absolute H, first row-major maximum, equality traceback) and (b) a CPU model of
 * k_fill1_v3 + k_trace (centred dot-product exponent in fp32, difference-form recurrence, sign-bit codes), then walks
 * model (b)'s path and records the smallest decision margins met on it.
 *
 *   gcc -O2 -fopenmp -ffp-contract=off -o tools/tie_study tools/tie_study.c -lm
 *   tools/tie_study chains.bin n_pairs seed
 * chains.bin: int64 N, int64 d, int64 offsets[N+1], double coords[total*3], double tensors[total*d]
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CSEG 10
#ifndef FORCE_STOP
#define FORCE_STOP 1
#endif
static inline float f_ex2(float x) { float y = exp2f(x); return (fabsf(y) < 1.17549435e-38f) ? 0.f : y; }

typedef struct {
    int mismatch, start_mismatch;
    double min_rel;      /* min over path cells of (d - second)/d */
    double min_rel_s;    /* same, scale = max(d, local neighbours' d) */
    double min_d;        /* min over path cells of d */
    double min_abs;      /* min over path cells of d - second */
    double start_grow;   /* smallest positive growth of H[i][m] at rows >= i* (fp32), or after the istar with threshold */
    double sc32_exact;   /* fp64-exact score of the fp32 path minus the reference score, relative */
    int ncols_diff;
    int first_div_kind;
    double div_s, div_a, div_b, div_H;
    double m1, m2, m3, mstart;   /* see VISIT */
    double v2, v2rel, st2[8];
    double v5, v6, v7, v8;
    unsigned char vg[8][4];
    unsigned char vs[4];   /* vg[3][2] OR sym margin2 < k 2^-24 * max d over the lane segment in this and the previous row, k = 2, 4, 8, 16 */   /* proposed flag: margin to a higher-priority candidate (walk-left stop cell: to left) < c 2^-53 Hseg + eps d */   /* v7: non-diag path cell, min d over its 10-cell segment / Hseg; v8: same with 5-cell half segments */   /* v5: min over non-diag path cells of d/Hseg; v6: min over non-diag path cells of (d - best higher-priority)/d */
} Diag;

static void study_pair(const double *t1, int n, const double *t2, int m, int d, const double *mean, double gamma, Diag *out)
{
    const int W = m + 1;
    double *S64 = malloc(sizeof(double) * n * m);
    double *H = calloc((size_t)(n + 1) * W, sizeof(double));
    float *S32 = malloc(sizeof(float) * n * m);
    float *A32 = malloc(sizeof(float) * (size_t)(n + 1) * W), *B32 = malloc(sizeof(float) * (size_t)(n + 1) * W), *D32 = malloc(sizeof(float) * (size_t)(n + 1) * W);
    uint8_t *code = calloc((size_t)(n + 1) * W, 1);
    /* reference */
    for (int a = 0; a < n; ++a)
        for (int b = 0; b < m; ++b) {
            double acc = 0.0;
            for (int k = 0; k < d; ++k) { double t = t1[a * d + k] - t2[b * d + k]; acc = acc + t * t; }
            S64[a * m + b] = exp(-gamma * acc);
        }
    for (int i = 1; i <= n; ++i)
        for (int j = 1; j <= m; ++j) {
            double dg = H[(i - 1) * W + j - 1] + S64[(i - 1) * m + j - 1], lf = H[i * W + j - 1], up = H[(i - 1) * W + j];
            double h = dg > lf ? dg : lf; h = h > up ? h : up; if (h < 0) h = 0;
            H[i * W + j] = h;
        }
    int ri = -1, rj = -1; double best = 0;
    for (int i = 1; i <= n; ++i)
        for (int j = 1; j <= m; ++j)
            if (H[i * W + j] > best) { best = H[i * W + j]; ri = i; rj = j; }
    /* fp32 model */
    const double g2 = gamma * 1.4426950408889634, sc = sqrt(2.0 * g2);
    float *r1 = malloc(sizeof(float) * n * (d + 1)), *r2 = malloc(sizeof(float) * m * (d + 1));
    for (int a = 0; a < n; ++a) { double nn = 0; for (int k = 0; k < d; ++k) { double x = t1[a * d + k] - mean[k]; nn += x * x; r1[a * (d + 1) + k] = (float)(sc * x); } r1[a * (d + 1) + d] = (float)(-g2 * nn); }
    for (int b = 0; b < m; ++b) { double nn = 0; for (int k = 0; k < d; ++k) { double x = t2[b * d + k] - mean[k]; nn += x * x; r2[b * (d + 1) + k] = (float)(sc * x); } r2[b * (d + 1) + d] = (float)(-g2 * nn); }
    for (int a = 0; a < n; ++a)
        for (int b = 0; b < m; ++b) {
            float e = r2[b * (d + 1) + d];
            for (int k = 0; k < d; ++k) e = fmaf(r1[a * (d + 1) + k], r2[b * (d + 1) + k], e);
            e = e + r1[a * (d + 1) + d];
            S32[a * m + b] = f_ex2(e);
        }
    float *u = calloc(W, sizeof(float));
    int istar = -1;
    float *grow = calloc(n + 1, sizeof(float));
    for (int i = 1; i <= n; ++i) {
        float a = 0.f;
        for (int j = 1; j <= m; ++j) {
            float s = S32[(i - 1) * m + j - 1], b = u[j];
            float dd = s > a ? s : a; dd = dd > b ? dd : b;
            code[i * W + j] = (dd == s) ? 1 : (dd == a ? 2 : 3);
            A32[i * W + j] = a; B32[i * W + j] = b; D32[i * W + j] = dd;
            u[j] = dd - a;
            a = dd - b;
        }
        grow[i] = a;
        if (a > 0.f) istar = i;
    }
    memset(out, 0, sizeof(*out));
    out->min_rel = out->min_rel_s = out->min_d = out->min_abs = out->start_grow = 1e300;
    out->m1 = out->m2 = out->m3 = out->mstart = out->v2 = out->v2rel = out->v5 = out->v6 = out->v7 = out->v8 = 1e300;
    if (ri < 0 || istar < 0) { out->mismatch = (ri < 0) != (istar < 0); goto done; }
    /* margins helper */
#define VISIT(i, j) VISITF(i, j, 0)
#define VISITF(i, j, FORCE_A) do { \
        float s_ = S32[((i) - 1) * m + (j) - 1], a_ = A32[(i) * W + (j)], b_ = B32[(i) * W + (j)], d_ = D32[(i) * W + (j)]; \
        float lo_ = s_ < a_ ? s_ : a_; lo_ = lo_ < b_ ? lo_ : b_; \
        double sec_ = ((double)s_ + a_ + b_) - d_ - lo_; \
        double mg_ = (double)d_ - sec_; \
        double scale_ = d_; \
        if ((j) > 1 && D32[(i) * W + (j) - 1] > scale_) scale_ = D32[(i) * W + (j) - 1]; \
        if ((i) > 1 && D32[((i) - 1) * W + (j)] > scale_) scale_ = D32[((i) - 1) * W + (j)]; \
        if (d_ > 0) { if (mg_ / d_ < out->min_rel) out->min_rel = mg_ / d_; if (mg_ / scale_ < out->min_rel_s) out->min_rel_s = mg_ / scale_; } \
        if (d_ < out->min_d) out->min_d = d_; \
        if (mg_ < out->min_abs) out->min_abs = mg_; \
        { const int seg_ = ((j) - 1) / CSEG; int ce_ = (seg_ + 1) * CSEG; if (ce_ > m) ce_ = m; \
          double hs_ = H[(i) * W + ce_]; if (hs_ <= 0) hs_ = 1e-300; \
          double hc_ = H[(i) * W + (j)]; if (hc_ <= 0) hc_ = 1e-300; \
          if (mg_ / hc_ < out->m1) out->m1 = mg_ / hc_; \
          if (d_ / hs_ < out->m3) out->m3 = d_ / hs_; \
          float sm_ = 1e30f; for (int jj_ = seg_ * CSEG + 1; jj_ <= ce_; ++jj_) if (D32[(i) * W + jj_] < sm_) sm_ = D32[(i) * W + jj_]; \
          if (sm_ / hs_ < out->m2) out->m2 = sm_ / hs_; \
          { double hp_ = 1e300; /* smallest margin to a HIGHER-priority candidate */ \
            if (d_ != s_) { hp_ = (double)d_ - s_; if (d_ != a_ && (double)d_ - a_ < hp_) hp_ = (double)d_ - a_; } \
            if (d_ != s_) { if (d_ / hs_ < out->v5) out->v5 = d_ / hs_; if (sm_ / hs_ < out->v7) out->v7 = sm_ / hs_; \
              { const int h0_ = ((j) - 1) / 5 * 5 + 1; float sm5_ = 1e30f; for (int jj_ = h0_; jj_ < h0_ + 5 && jj_ <= m; ++jj_) if (D32[(i) * W + jj_] < sm5_) sm5_ = D32[(i) * W + jj_]; \
                if (sm5_ / hs_ < out->v8) out->v8 = sm5_ / hs_; } } \
            { static const double cs__[8] = {1, 4, 16, 64, 256, 1024, 4096, 65536}; static const double es__[4] = {0, 1e-5, 1e-4, 1e-3}; \
              double zm_ = hp_; if (FORCE_A && (double)d_ - a_ > 0 && (double)d_ - a_ < zm_) zm_ = (double)d_ - a_; \
              for (int ci_ = 0; ci_ < 8; ++ci_) for (int ei_ = 0; ei_ < 4; ++ei_) if (zm_ < cs__[ci_] * 1.1102230246251565e-16 * hs_ + es__[ei_] * d_) out->vg[ci_][ei_] = 1; } \
            { float dm_ = 0.f; for (int jj_ = seg_ * CSEG + 1; jj_ <= ce_; ++jj_) { if (D32[(i) * W + jj_] > dm_) dm_ = D32[(i) * W + jj_]; if ((i) > 1 && D32[((i) - 1) * W + jj_] > dm_) dm_ = D32[((i) - 1) * W + jj_]; } \
              static const double ks__[4] = {2, 4, 8, 16}; \
              for (int k_ = 0; k_ < 4; ++k_) if (mg_ < ks__[k_] * 5.9604644775390625e-08 * dm_ && mg_ < 0.5 * d_) out->vs[k_] = 1; } \
            if (hp_ / hc_ < out->v2) out->v2 = hp_ / hc_; if (d_ > 0 && hp_ / d_ < out->v2rel) out->v2rel = hp_ / d_; } } \
    } while (0)
    for (int i = 1; i <= n; ++i) if (grow[i] > 0.f && grow[i] < out->start_grow) out->start_grow = grow[i];
    { /* mstart: largest relative growth among the rows after the last row whose growth is >= 2^-40 H (what a thresholded istar would skip) */
      for (int i = 1; i <= n; ++i) if (grow[i] > 0.f) { double r = grow[i] / (H[i * W + m] > 0 ? H[i * W + m] : 1e-300); if (r < out->mstart) out->mstart = r; } }
    { const double cs_[8] = {1, 4, 16, 64, 256, 1024, 4096, 65536};
      for (int k = 0; k < 8; ++k) { int sig = 0; for (int i = 1; i <= n; ++i) if (grow[i] / (H[i * W + m] > 0 ? H[i * W + m] : 1e-300) >= cs_[k] * 1.1102230246251565e-16) sig = i;
        out->st2[k] = (istar != sig); } }
    int bi = istar, bj = m;
    while (bj > 1 && code[bi * W + bj] == 2) { VISIT(bi, bj); --bj; }
    if (bj > 1) VISITF(bi, bj, FORCE_STOP);
    out->start_mismatch = (bi != ri || bj != rj);
    /* walk both */
    {
        int i = bi, j = bj, i2 = ri, j2 = rj;
        int n32 = 0, n64 = 0;
        int *p32 = malloc(sizeof(int) * 2 * (n + m + 2)), *p64 = malloc(sizeof(int) * 2 * (n + m + 2));
        double ex = 0.0;
        while (i > 0 && j > 0) {
            VISIT(i, j);
            int cd = code[i * W + j];
            if (D32[i * W + j] == 0.f) break;
            if (cd == 1) { --i; --j; p32[2 * n32] = i; p32[2 * n32 + 1] = j; ++n32; ex += S64[i * m + j]; }
            else if (cd == 2) { --j; p32[2 * n32] = -1; p32[2 * n32 + 1] = j; ++n32; }
            else { --i; p32[2 * n32] = i; p32[2 * n32 + 1] = -1; ++n32; }
        }
        i = i2; j = j2;
        while (i > 0 && j > 0) {
            double h = H[i * W + j];
            if (h == 0) break;
            if (h == H[(i - 1) * W + j - 1] + S64[(i - 1) * m + j - 1]) { --i; --j; p64[2 * n64] = i; p64[2 * n64 + 1] = j; ++n64; }
            else if (h == H[i * W + j - 1]) { --j; p64[2 * n64] = -1; p64[2 * n64 + 1] = j; ++n64; }
            else { --i; p64[2 * n64] = i; p64[2 * n64 + 1] = -1; ++n64; }
        }
        out->mismatch = (n32 != n64) || memcmp(p32, p64, sizeof(int) * 2 * n32) != 0;
        /* aligned-column difference */
        if (out->mismatch) {
            int cnt = 0;
            for (int q = 0; q < n32; ++q) if (p32[2 * q] >= 0 && p32[2 * q + 1] >= 0) {
                int found = 0;
                for (int r = 0; r < n64; ++r) if (p64[2 * r] == p32[2 * q] && p64[2 * r + 1] == p32[2 * q + 1]) { found = 1; break; }
                cnt += !found;
            }
            out->ncols_diff = cnt;
            /* first divergence (from the end): cell where the two walks, started equal, take different moves */
            if (!out->start_mismatch) {
                int ii = bi, jj = bj;
                for (int q = 0; q < n32 && q < n64; ++q) {
                    if (p32[2 * q] != p64[2 * q] || p32[2 * q + 1] != p64[2 * q + 1]) {
                        out->first_div_kind = code[ii * W + jj];
                        out->div_s = S32[(ii - 1) * m + jj - 1]; out->div_a = A32[ii * W + jj]; out->div_b = B32[ii * W + jj];
                        out->div_H = H[ii * W + jj];
                        break;
                    }
                    int cd = code[ii * W + jj];
                    if (cd == 1) { --ii; --jj; } else if (cd == 2) --jj; else --ii;
                }
            } else {
                out->first_div_kind = -1;
                out->div_H = best;
            }
        }
        out->sc32_exact = best > 0 ? (ex - best) / best : 0;
        free(p32); free(p64);
    }
done:
    free(S64); free(H); free(S32); free(A32); free(B32); free(D32); free(code); free(r1); free(r2); free(u); free(grow);
}

int main(int argc, char **argv)
{
    if (argc < 4) { fprintf(stderr, "usage: tie_study chains.bin n_pairs seed [unrelated_only]\n"); return 1; }
    FILE *f = fopen(argv[1], "rb");
    int64_t N, d;
    if (!f || fread(&N, 8, 1, f) != 1 || fread(&d, 8, 1, f) != 1) return 2;
    int64_t *off = malloc(8 * (N + 1));
    if (fread(off, 8, N + 1, f) != (size_t)(N + 1)) return 2;
    const int64_t total = off[N];
    double *coords = malloc(8 * total * 3), *tens = malloc(8 * total * d);
    if (fread(coords, 8, total * 3, f) != (size_t)(total * 3) || fread(tens, 8, total * d, f) != (size_t)(total * d)) return 2;
    fclose(f);
    double mean[32] = {0};
    for (int64_t r = 0; r < total; ++r) for (int k = 0; k < d; ++k) mean[k] += tens[r * d + k];
    for (int k = 0; k < d; ++k) mean[k] /= (double)total;
    int np = atoi(argv[2]);
    srand48(atol(argv[3]));
    int *pi = malloc(4 * (np + 1)), *pj = malloc(4 * (np + 1));
    if (argc >= 6) { np = 1; pi[0] = atoi(argv[4]); pj[0] = atoi(argv[5]); }      /* one given pair: tie_study chains.bin 1 0 i j */
    else for (int q = 0; q < np; ++q) {
        int i = (int)(drand48() * N), j = (int)(drand48() * N);
        if (i == j) { --q; continue; }
        if (i > j) { int t = i; i = j; j = t; }
        pi[q] = i; pj[q] = j;
    }
    Diag *dg = malloc(sizeof(Diag) * np);
#pragma omp parallel for schedule(dynamic, 8)
    for (int q = 0; q < np; ++q)
        study_pair(tens + off[pi[q]] * d, (int)(off[pi[q] + 1] - off[pi[q]]), tens + off[pj[q]] * d, (int)(off[pj[q] + 1] - off[pj[q]]), (int)d, mean, 7.0, &dg[q]);
    int nm = 0;
    for (int q = 0; q < np; ++q) nm += dg[q].mismatch;
    printf("pairs %d mismatches %d (%.3f %%)\n", np, nm, 100.0 * nm / np);
    printf("# mismatching pairs: i j start_mm kind ncols_diff min_rel min_rel_s min_abs min_d start_grow sc32_exact_rel div(s,a,b,H)\n");
    for (int q = 0; q < np; ++q) if (dg[q].mismatch)
        printf("MM%s %d %d %d %d %d %.3e %.3e %.3e %.3e %.3e %.3e | %.3e %.3e %.3e %.4g\n", (dg[q].st2[3] || dg[q].vg[3][2]) ? "" : "-MISSED", pi[q], pj[q], dg[q].start_mismatch, dg[q].first_div_kind, dg[q].ncols_diff,
               dg[q].min_rel, dg[q].min_rel_s, dg[q].min_abs, dg[q].min_d, dg[q].start_grow, dg[q].sc32_exact, dg[q].div_s, dg[q].div_a, dg[q].div_b, dg[q].div_H);
    /* flag rates for candidate criteria over all pairs */
    const double rels[] = {1e-3, 1e-4, 1e-5, 1e-6, 1e-7, 0};
    const double abss[] = {1e-8, 1e-10, 1e-12, 1e-13, 1e-14, 1e-16, 0};
    printf("# criterion: on-path min_rel < R or min_d < T or start_grow < T   -> flagged pairs, missed mismatches\n");
    for (int a = 0; a < 6; ++a)
        for (int b = 0; b < 7; ++b) {
            int fl = 0, missed = 0;
            for (int q = 0; q < np; ++q) {
                int flag = (dg[q].min_rel <= rels[a]) || (dg[q].min_d < abss[b]) || (dg[q].start_grow < abss[b]);
                fl += flag;
                if (dg[q].mismatch && !flag) ++missed;
            }
            printf("R=%.0e T=%.0e flagged %d (%.3f %%) missed %d\n", rels[a], abss[b], fl, 100.0 * fl / np, missed);
        }
    printf("# same with min_rel_s (margin relative to the largest d of the cell and its left/up neighbours)\n");
    for (int a = 0; a < 6; ++a)
        for (int b = 0; b < 7; ++b) {
            int fl = 0, missed = 0;
            for (int q = 0; q < np; ++q) {
                int flag = (dg[q].min_rel_s <= rels[a]) || (dg[q].min_d < abss[b]) || (dg[q].start_grow < abss[b]);
                fl += flag;
                if (dg[q].mismatch && !flag) ++missed;
            }
            printf("RS=%.0e T=%.0e flagged %d (%.3f %%) missed %d\n", rels[a], abss[b], fl, 100.0 * fl / np, missed);
        }
    {
        const double cs[] = {1, 4, 16, 64, 256, 1024, 4096, 65536};
        const double u = 1.1102230246251565e-16;
        printf("# ideal: on-path cell margin2 / H < c 2^-53 (m1) | cell d / Hseg (m3) | segment min d / Hseg (m2); all OR start growth rel < c 2^-53; class-2 separately RS<1e-6\n");
        for (int b = 0; b < 8; ++b) {
            int f1 = 0, f2 = 0, f3 = 0, x1 = 0, x2 = 0, x3 = 0;
            for (int q = 0; q < np; ++q) {
                const int st = dg[q].mstart < cs[b] * u, c2 = dg[q].min_rel <= 1e-4;
                const int a1 = st || c2 || dg[q].m1 < cs[b] * u, a2 = st || c2 || dg[q].m2 < cs[b] * u, a3 = st || c2 || dg[q].m3 < cs[b] * u;
                f1 += a1; f2 += a2; f3 += a3;
                if (dg[q].mismatch) { x1 += !a1; x2 += !a2; x3 += !a3; }
            }
            printf("c=%6.0f  m1: flagged %.3f %% missed %d | m3(cell d): %.3f %% missed %d | m2(segment d): %.3f %% missed %d\n", cs[b], 100.0 * f1 / np, x1, 100.0 * f3 / np, x3, 100.0 * f2 / np, x2);
        }
        int fs = 0, fc2 = 0; for (int q = 0; q < np; ++q) { fs += dg[q].mstart < 64 * u; fc2 += dg[q].min_rel <= 1e-4; }
        printf("start-only flagged (c=64) %.3f %%, class-2-only flagged (RS<=1e-6) %.3f %%\n", 100.0 * fs / np, 100.0 * fc2 / np);
    }
    {
        const double cs[] = {1, 4, 16, 64, 256, 1024, 4096, 65536};
        const double u = 1.1102230246251565e-16;
        printf("# V2: on-path cell, margin to a higher-priority candidate / H < c 2^-53, OR thresholded start row differs (st2), [OR v2rel<=1e-4]\n");
        for (int b = 0; b < 8; ++b) {
            int f1 = 0, x1 = 0, fs = 0, f2 = 0, x2 = 0;
            for (int q = 0; q < np; ++q) {
                const int st = dg[q].st2[b] != 0;
                const int a1 = st || dg[q].v2 < cs[b] * u;
                const int a2 = a1 || dg[q].v2rel <= 1e-4;
                f1 += a1; fs += st; f2 += a2;
                if (dg[q].mismatch) { x1 += !a1; x2 += !a2; }
            }
            printf("c=%6.0f  V2|st2: flagged %.3f %% missed %d   (st2 alone %.3f %%)   with rel: %.3f %% missed %d\n", cs[b], 100.0 * f1 / np, x1, 100.0 * fs / np, 100.0 * f2 / np, x2);
        }
    }
    {
        const double cs[] = {1, 4, 16, 64, 256, 1024, 4096, 65536};
        const double u = 1.1102230246251565e-16;
        printf("# V5: on-path non-diag cell with d / Hseg < c 2^-53, OR st2, OR v2rel <= 1e-4\n");
        for (int b = 0; b < 8; ++b) {
            int f1 = 0, x1 = 0;
            for (int q = 0; q < np; ++q) {
                const int a1 = dg[q].st2[b] != 0 || dg[q].v5 < cs[b] * u || dg[q].v2rel <= 1e-4;
                f1 += a1;
                if (dg[q].mismatch) x1 += !a1;
            }
            printf("c=%6.0f  V5: flagged %.3f %% missed %d\n", cs[b], 100.0 * f1 / np, x1);
        }
    }
    {
        const double cs[] = {1, 4, 16, 64, 256, 1024, 4096, 65536};
        const double u = 1.1102230246251565e-16;
        printf("# V7/V8: on-path non-diag cell whose 10-cell (5-cell) segment holds a cell with d / Hseg < c 2^-53, OR st2, OR v2rel <= 1e-4\n");
        for (int b = 0; b < 8; ++b) {
            int f1 = 0, x1 = 0, f2 = 0, x2 = 0;
            for (int q = 0; q < np; ++q) {
                const int a1 = dg[q].st2[b] != 0 || dg[q].v7 < cs[b] * u || dg[q].v2rel <= 1e-4;
                const int a2 = dg[q].st2[b] != 0 || dg[q].v8 < cs[b] * u || dg[q].v2rel <= 1e-4;
                f1 += a1; f2 += a2;
                if (dg[q].mismatch) { x1 += !a1; x2 += !a2; }
            }
            printf("c=%6.0f  V7: flagged %.3f %% missed %d   V8: flagged %.3f %% missed %d\n", cs[b], 100.0 * f1 / np, x1, 100.0 * f2 / np, x2);
        }
    }
    {
        const double cs[] = {1, 4, 16, 64, 256, 1024, 4096, 65536};
        const double es[] = {0, 1e-5, 1e-4, 1e-3};
        printf("# PROPOSED: flag = st2(c) OR on-path cell with margin-to-higher-priority < c 2^-53 Hseg + eps d\n");
        for (int b = 0; b < 8; ++b) for (int e = 0; e < 4; ++e) {
            int f1 = 0, x1 = 0;
            for (int q = 0; q < np; ++q) {
                const int a1 = dg[q].st2[b] != 0 || dg[q].vg[b][e];
                f1 += a1;
                if (dg[q].mismatch) x1 += !a1;
            }
            printf("c=%6.0f eps=%.0e : flagged %.3f %% missed %d\n", cs[b], es[e], 100.0 * f1 / np, x1);
        }
    }
    {
        printf("# SCALE: st2(c=8) OR vg(c=8.., eps=1e-4) OR sym margin2 < k 2^-24 max(d over lane segment, this + previous row)\n");
        for (int k = 0; k < 4; ++k) {
            int f1 = 0, x1 = 0, f0 = 0;
            for (int q = 0; q < np; ++q) {
                const int base = dg[q].st2[2] != 0 || dg[q].vg[2][2];
                const int a1 = base || dg[q].vs[k];
                f1 += a1; f0 += base;
                if (dg[q].mismatch) x1 += !a1;
            }
            printf("k=%d : flagged %.3f %% (base %.3f %%) missed %d\n", 2 << k, 100.0 * f1 / np, 100.0 * f0 / np, x1);
        }
    }
    printf("# absolute margin only: min_abs < T or start_grow < T\n");
    for (int b = 0; b < 6; ++b) {
        int fl = 0, missed = 0;
        for (int q = 0; q < np; ++q) {
            int flag = (dg[q].min_abs < abss[b]) || (dg[q].start_grow < abss[b]);
            fl += flag;
            if (dg[q].mismatch && !flag) ++missed;
        }
        printf("A=%.0e flagged %d (%.3f %%) missed %d\n", abss[b], fl, 100.0 * fl / np, missed);
    }
    return 0;
}
