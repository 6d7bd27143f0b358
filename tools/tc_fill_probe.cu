// tc_fill_probe.cu -- steady-state and correctness probe of k_fill1_tc (crt_fill_tc.cuh): synthetic chains, rounds of 128 partners.
//   * correctness: H[n][m] of sampled pairs against a float64 Smith-Waterman on the host, and a walk over the GPU's traceback
//     codes (sum of S over the diagonal moves == H[n][m]: the codes describe an optimal path, so the layout is decoded right)
//   * exponent accuracy of the tensor-core tile (via S = 2^E on the path) is implied by the score check
//   * throughput: cells per second against k_fill1_v4's 1.05e12
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I../caretta_b200/csrc -o tc_fill_probe tc_fill_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <random>
#include <algorithm>
#include "crt_fill_tc.cuh"

using namespace crt;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

int main(int argc, char **argv)
{
    const int L = argc > 1 ? atoi(argv[1]) : 300;            // chain length (columns and rows)
    const int waves = argc > 2 ? atoi(argv[2]) : 4;          // rounds = 148 * 2 * waves
    const int ragged = argc > 3 ? atoi(argv[3]) : 0;         // 1: partner lengths L - (lane % 37)
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("device %s, %d SMs, %d MHz\n", pr.name, pr.multiProcessorCount, clk_khz / 1000);
    const int RS = 12, NCH = 256;                            // 256 synthetic chains of length L
    const long long total = (long long)NCH * L;
    std::mt19937 rng(3); std::normal_distribution<double> nd(0.0, 0.3);
    const double g2 = 7.0 * 1.4426950408889634, sc = std::sqrt(2.0 * g2);
    std::vector<float> rec((size_t)(total + 96) * RS, 0.f);
    for (long long r = 0; r < total; ++r) {
        float *o = rec.data() + (size_t)(r + 48) * RS; double nn = 0;
        // chains come in families so that pairs have real alignments: family = chain % 8, residue noise on a family template
        std::mt19937 fr((unsigned)((r / L) % 8) * 7919u + (unsigned)(r % L));
        std::normal_distribution<double> fd(0.0, 0.3);
        for (int k = 0; k < 10; ++k) { const double x = fd(fr) + 0.05 * nd(rng); nn += x * x; o[k] = (float)(sc * x); }
        o[10] = (float)(-g2 * nn); o[11] = 1.f;
    }
    const int n_rounds = pr.multiProcessorCount * 2 * waves;
    const int n_strips = (L + TC_SC - 1) / TC_SC;
    const int strip_w = ((L + n_strips - 1) / n_strips + 15) / 16 * 16;
    int tiles_per_row = 0;
    for (int s = 0; s < n_strips; ++s) { const int w = std::min(strip_w, ((L - s * strip_w + 15) / 16) * 16); tiles_per_row += (w + 31) / 32; }
    std::vector<TcRound> rounds(n_rounds);
    std::vector<TcPartner> parts((size_t)n_rounds * TC_LANES);
    long long tb_n = 0;
    for (int r = 0; r < n_rounds; ++r) {
        TcRound &R = rounds[r];
        const int j = r % NCH;
        R.bnd_base = (long long)r * L * TC_LANES; R.col_base = j * L; R.col_chain = j; R.m = L; R.n_strips = n_strips; R.strip_w = strip_w;
        R.part_base = r * TC_LANES; R.n_part = TC_LANES; R.max_rows = L;
        for (int l = 0; l < TC_LANES; ++l) {
            TcPartner &P = parts[(size_t)r * TC_LANES + l];
            const int i = (j + 1 + l) % NCH;
            P.row_base = i * L; P.row_chain = i; P.n = ragged ? L - (l % 37) : L; P.slot = r * TC_LANES + l; P.round = r;
            P.tb_base = tb_n; tb_n += (long long)P.n * tiles_per_row;
        }
    }
    printf("L = %d: %d strips of %d columns, %d tiles per row, %d rounds, traceback %.2f GB (%.3f B per cell)\n", L, n_strips, strip_w, tiles_per_row,
           n_rounds, tb_n * 16e-9, tb_n * 16.0 / ((double)n_rounds * TC_LANES * L * L));
    float *d_rec, *d_bnd; int *d_counter; TcRound *d_rounds; TcPartner *d_parts; uint4 *d_tb; int *d_istar, *d_zflag; double *d_score;
    const size_t np = (size_t)n_rounds * TC_LANES;
    CK(cudaMalloc(&d_rec, rec.size() * 4)); CK(cudaMemcpy(d_rec, rec.data(), rec.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_bnd, (size_t)n_rounds * L * TC_LANES * 4 + 16));
    CK(cudaMalloc(&d_rounds, rounds.size() * sizeof(TcRound))); CK(cudaMemcpy(d_rounds, rounds.data(), rounds.size() * sizeof(TcRound), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_parts, parts.size() * sizeof(TcPartner))); CK(cudaMemcpy(d_parts, parts.data(), parts.size() * sizeof(TcPartner), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_tb, (size_t)tb_n * 16 + 16)); CK(cudaMemset(d_tb, 0xff, (size_t)tb_n * 16));
    CK(cudaMalloc(&d_istar, np * 4)); CK(cudaMalloc(&d_zflag, np * 4)); CK(cudaMalloc(&d_score, np * 8));
    CK(cudaMalloc(&d_counter, 4));
    TcFill1Args a{};
    a.rec = d_rec + 48 * RS; a.rounds = d_rounds; a.partners = d_parts; a.tb = d_tb; a.bnd = d_bnd; a.pair_istar = d_istar; a.pair_zflag = d_zflag;
    a.pair_score = d_score; a.counter = d_counter;
    long long *d_prof; CK(cudaMalloc(&d_prof, 64)); CK(cudaMemset(d_prof, 0, 64)); a.prof = d_prof; a.tie = TieArgs{(float)(8.0 * 1.1102230246251565e-16), 1e-4f};
    const size_t smem = (size_t)(2 * TC_LANES + TC_SC) * 48 * 4;
    CK(cudaFuncSetAttribute(k_fill1_tc<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_fill1_tc<12>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_fill1_tc<12>, TC_THREADS, smem));
    printf("dynamic smem %zu B, occupancy %d CTAs/SM\n", smem, occ);
    {
        cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, k_fill1_tc<12>));
        printf("regs %d, static smem %zu, maxDyn %d, carveout %d, smem/SM %zu, smem/block optin %zu, regs/SM %d\n", fa.numRegs, fa.sharedSizeBytes, fa.maxDynamicSharedSizeBytes,
               fa.preferredShmemCarveout, pr.sharedMemPerMultiprocessor, pr.sharedMemPerBlockOptin, pr.regsPerMultiprocessor);
        for (int bs : {32, 64, 96, 128, 160, 192, 256}) { int o = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k_fill1_tc<12>, bs, 0); printf("  block %d -> occ %d\n", bs, o); }
        for (size_t sm : {0ul, 16384ul, 32768ul, 49152ul, 66560ul}) { int o = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k_fill1_tc<12>, TC_THREADS, sm); printf("  dyn %zu -> occ %d\n", sm, o); }
    }
    const int per_sm = argc > 4 ? atoi(argv[4]) : 2;
    const int grid = std::min(n_rounds, per_sm * pr.multiProcessorCount);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaMemsetAsync(d_counter, 0, 4)); k_fill1_tc<12><<<grid, TC_THREADS, smem>>>(a, n_rounds);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        CK(cudaMemsetAsync(d_counter, 0, 4)); k_fill1_tc<12><<<grid, TC_THREADS, smem>>>(a, n_rounds);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = std::min(best, ms);
    }
    double cells = 0;
    for (auto &P : parts) cells += (double)P.n * L;
    const double cyc_cell = best * 1e-3 * clk_khz * 1e3 * pr.multiProcessorCount * 4 / (cells / 32);
    printf("k_fill1_tc: %.3f ms, %.1f Gcell/s (k_fill1_v4: 1050), %.2f cycles per warp-cell per SMSP\n", best, cells / (best * 1e-3) / 1e9, cyc_cell);

    {
        long long hp[4]; CK(cudaMemcpy(hp, d_prof, 32, cudaMemcpyDeviceToHost));
        if (hp[3]) printf("profile of DP warp 0 of CTA 0: %.0f cycles per tile in total, %.0f waiting for the exponent tile, %.0f in the per-row code (per tile)\n", (double)hp[2] / hp[3], (double)hp[0] / hp[3], (double)hp[1] / hp[3]);
    }
    // ---- correctness on sampled pairs
    std::vector<double> score(np); std::vector<int> istar(np);
    CK(cudaMemcpy(score.data(), d_score, np * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(istar.data(), d_istar, np * 4, cudaMemcpyDeviceToHost));
    int bad_score = 0, bad_walk = 0, checked = 0; double max_rel = 0, max_walk = 0;
    std::vector<uint4> tb;
    for (size_t p = 0; p < np; p += np / 97 + 1) {
        const TcPartner &P = parts[p]; const TcRound &R = rounds[P.round];
        const int n = P.n, m = R.m;
        std::vector<double> S((size_t)n * m), H((size_t)(n + 1) * (m + 1), 0.0);
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < m; ++j) {
                const float *x = rec.data() + (size_t)(48 + P.row_base + i) * RS, *y = rec.data() + (size_t)(48 + R.col_base + j) * RS;
                double e = (double)x[10] + (double)y[10];
                for (int k = 0; k < 10; ++k) e += (double)x[k] * (double)y[k];
                S[(size_t)i * m + j] = std::exp2(e);
            }
        for (int i = 1; i <= n; ++i)
            for (int j = 1; j <= m; ++j)
                H[(size_t)i * (m + 1) + j] = std::max({H[(size_t)(i - 1) * (m + 1) + j - 1] + S[(size_t)(i - 1) * m + j - 1], H[(size_t)i * (m + 1) + j - 1], H[(size_t)(i - 1) * (m + 1) + j]});
        const double ref = H[(size_t)n * (m + 1) + m];
        const double rel = std::fabs(score[p] - ref) / std::max(ref, 1e-30);
        max_rel = std::max(max_rel, rel);
        if (!(rel < 1e-5)) { if (bad_score < 5) printf("  pair %zu: H[n][m] gpu %.9g ref %.9g\n", p, score[p], ref); ++bad_score; }
        // walk the codes
        const size_t words = (size_t)n * tiles_per_row;
        tb.resize(words);
        CK(cudaMemcpy(tb.data(), d_tb + P.tb_base, words * 16, cudaMemcpyDeviceToHost));
        const int tiles_full = (R.strip_w + 31) / 32;
        auto code = [&](int i, int j) -> unsigned {           // 0-based cell
            const int strip = j / R.strip_w, c = j - strip * R.strip_w;
            const int w_cols = std::min(R.strip_w, ((m - strip * R.strip_w + 15) / 16) * 16), nt = (w_cols + 31) / 32;
            const uint4 v = tb[(size_t)strip * n * tiles_full + (size_t)i * nt + (c >> 5)];
            const unsigned ws[3] = {v.x, v.y, v.z};
            unsigned out = 0;
            for (int k = 0; k < 3; ++k) { const int pbit = 3 * (c & 31) + k; out = (out << 1) | ((ws[pbit >> 5] >> (31 - (pbit & 31))) & 1u); }
            return out;                                        // (S attains) << 2 | (left attains) << 1 | suspect
        };
        int i = istar[p] & ~ISTAR_TIE, j = m;
        double walk = 0; int steps = 0;
        if (i > 0) {
            while (j > 1 && (code(i - 1, j - 1) & 2u)) --j;      // first column of row i* that attains the maximum
            while (i > 0 && j > 0 && steps < n + m + 5) {
                const unsigned cd = code(i - 1, j - 1);
                if (cd & 4u) { walk += S[(size_t)(i - 1) * m + j - 1]; --i; --j; }
                else if (cd & 2u) --j;
                else --i;
                ++steps;
            }
        }
        const double wrel = std::fabs(walk - ref) / std::max(ref, 1e-30);
        max_walk = std::max(max_walk, wrel);
        if (!(wrel < 1e-5)) { if (bad_walk < 5) printf("  pair %zu: walk %.9g ref %.9g (istar %d)\n", p, walk, ref, istar[p] & ~ISTAR_TIE); ++bad_walk; }
        ++checked;
    }
    printf("checked %d pairs: H[n][m] max rel err %.2e (%d outside 1e-5); code walk max rel err %.2e (%d outside 1e-5)\n", checked, max_rel, bad_score, max_walk, bad_walk);
    return bad_score || bad_walk;
}
