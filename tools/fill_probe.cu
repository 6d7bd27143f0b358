// Steady-state probe of the two fp32 fill kernels (sm_100a): identical synthetic units, grid = k resident warps per SM,
// reports cycles per row-step per warp and per SMSP.  Separates per-warp latency from pipe throughput and from the
// batch-tail effects of the real run.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../caretta_b200/csrc -o fill_probe fill_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <random>
#include <cstring>
#include "crt_fill_f32.cuh"
#ifdef PROBE_V2
#include "crt_fill1_v2.cuh"
#include "crt_fill1_v4.cuh"
#include "crt_fill2_v3.cuh"
#endif

using namespace crt;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

static int g_sms = 148;
static double g_clk = 1.965e9;

struct Data {
    float *rec; int *meta; float4 *rows2, *cols2;
    uint4 *tb; int *istar, *zflag; double *score; Unit *units;
    int L, nchains, G;
};

template <typename F>
static float time_kernel(F launch, int reps = 3)
{
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

static void report(const char *name, int C, int nstrips, int n_units, const Data &d, float ms, int m)
{
    const double rowsteps = (double)n_units * (d.G + 31) * nstrips;          // warp row-steps
    const double cyc = ms * 1e-3 * g_clk;
    const double per_smsp = cyc * g_sms * 4 / rowsteps;                       // cycles per row-step per SMSP
    const double warps_per_smsp = (double)n_units / (g_sms * 4);
    const double cells = (double)n_units * d.G * m;
    printf("%-26s C=%2d strips=%d units=%5d (%.2f warps/SMSP) %8.3f ms  %7.1f cyc/rowstep/SMSP  %7.1f cyc/rowstep/warp  %6.1f Gcell/s\n",
           name, C, nstrips, n_units, warps_per_smsp, ms, per_smsp, per_smsp * (warps_per_smsp < 1 ? 1 : warps_per_smsp),
           cells / (ms * 1e-3) / 1e9);
}

int main(int argc, char **argv)
{
    const int L = argc > 1 ? atoi(argv[1]) : 300;
    const int chains_per_unit = argc > 2 ? atoi(argv[2]) : 10;
    const char *which = argc > 3 ? argv[3] : "all";          // all | v1 | v2 | v3 | c6 | f2
    const int only_k = argc > 4 ? atoi(argv[4]) : 0;         // 0 = sweep the resident-warp counts
    auto want = [&](const char *name) { return !strcmp(which, "all") || !strcmp(which, name); };
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    g_sms = pr.multiProcessorCount;
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0); g_clk = clk_khz * 1e3;
    printf("device %s, %d SMs, %.0f MHz\n", pr.name, g_sms, g_clk / 1e6);

    Data d{}; d.L = L; d.nchains = 64; d.G = L * chains_per_unit;
    const int RS = 12;
    const long long total = (long long)d.nchains * L;
    std::mt19937 rng(1); std::normal_distribution<float> nd(0.f, 1.f);
    std::vector<float> rec((size_t)(total + 2 * ROW_PAD) * RS, 0.f);
    std::vector<int> meta((size_t)total + 2 * ROW_PAD, 0);
    std::vector<float4> c2((size_t)total + 2 * ROW_PAD);
    for (long long r = 0; r < total; ++r) {
        float *o = rec.data() + (size_t)(r + ROW_PAD) * RS; float nn = 0;
        for (int k = 0; k < 10; ++k) { o[k] = 0.8f * nd(rng); nn += o[k] * o[k]; }
        o[10] = -0.5f * nn; o[11] = 1.f;
        const int ch = (int)(r / L), pos = (int)(r % L);
        meta[(size_t)r + ROW_PAD] = make_meta(ch, pos == 0, pos == L - 1);
        c2[(size_t)r + ROW_PAD] = make_float4(0.2f * pos + nd(rng), nd(rng), nd(rng), 0.f);
    }
    CK(cudaMalloc(&d.rec, rec.size() * 4)); CK(cudaMemcpy(d.rec, rec.data(), rec.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d.meta, meta.size() * 4)); CK(cudaMemcpy(d.meta, meta.data(), meta.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d.cols2, c2.size() * 16)); CK(cudaMemcpy(d.cols2, c2.data(), c2.size() * 16, cudaMemcpyHostToDevice));
    // stage-2 rows: (x,y,z,meta bits)
    std::vector<float4> r2((size_t)total + 2 * ROW_PAD);
    for (size_t q = 0; q < r2.size(); ++q) { r2[q] = c2[q]; memcpy(&r2[q].w, &meta[q], 4); }
    CK(cudaMalloc(&d.rows2, r2.size() * 16)); CK(cudaMemcpy(d.rows2, r2.data(), r2.size() * 16, cudaMemcpyHostToDevice));

    const int max_units = g_sms * 32;
    const int tchunks = (d.G + 31 + 3) / 4;
    const size_t tb_per_unit = (size_t)2 * tchunks * 32;          // room for 2 strips
    CK(cudaMalloc(&d.tb, (size_t)g_sms * 8 * 2 * tb_per_unit * 16));        // stage-1 runs at most 16 units per SM here
    CK(cudaMalloc(&d.istar, (size_t)max_units * 64 * 4)); CK(cudaMalloc(&d.zflag, (size_t)max_units * 64 * 4));
    CK(cudaMalloc(&d.score, (size_t)max_units * 64 * 8));
    float *bnd; CK(cudaMalloc(&bnd, (size_t)max_units * d.G * 4));
    CK(cudaMalloc(&d.units, (size_t)max_units * sizeof(Unit)));

    std::vector<long long> h_off(d.nchains + 1);
    for (int q = 0; q <= d.nchains; ++q) h_off[q] = (long long)q * L;
    long long *d_offsets; CK(cudaMalloc(&d_offsets, h_off.size() * 8)); CK(cudaMemcpy(d_offsets, h_off.data(), h_off.size() * 8, cudaMemcpyHostToDevice));
    auto make_units = [&](int n_units, int C, int nstrips) {
        std::vector<Unit> hu(n_units);
        for (int q = 0; q < n_units; ++q) {
            Unit u{};
            const int c0 = q % (d.nchains - chains_per_unit - 1);
            u.row_base = (long long)c0 * L; u.row_chain0 = c0; u.n_pairs = chains_per_unit; u.G = d.G;
            u.col_chain = d.nchains - 1; u.col_base = (d.nchains - 1) * L; u.m = L;
            u.tb_base = (long long)(q % (g_sms * 16)) * tb_per_unit; u.rows2_base = u.row_base; u.bnd_base = (long long)q * d.G;
            u.pair_base = q * chains_per_unit; u.n_strips = nstrips; u.tchunks = tchunks; u.path_stride = 2 * L;
            hu[q] = u;
        }
        CK(cudaMemcpy(d.units, hu.data(), sizeof(Unit) * n_units, cudaMemcpyHostToDevice));
    };
    FillOut fo{}; fo.tb = d.tb; fo.pair_istar = d.istar; fo.pair_zflag = d.zflag; fo.pair_score = d.score; fo.bnd = bnd;
    Fill1Args a1{d.rec + (size_t)ROW_PAD * RS, d.meta + ROW_PAD};
    Fill2Args a2{d.rows2 + ROW_PAD, d.cols2 + ROW_PAD};

    const int per_sm[] = {1, 2, 4, 6, 8, 9, 10, 11, 12, 16, 24, 32};
    if (want("v1")) for (int k : per_sm) {
        if (k > 16 || (only_k && k != only_k)) continue;
        const int n = g_sms * k;
        make_units(n, 10, 1);
        float ms = time_kernel([&] { k_fill1_f32<10, 10, false><<<n, 32>>>(d.units, n, a1, fo); });
        report("fill1<10,10>", 10, 1, n, d, ms, L);
    }
#ifdef PROBE_V2
    if (want("v2")) for (int k : per_sm) {
        if (k > 16 || (only_k && k != only_k)) continue;
        const int n = g_sms * k;
        make_units(n, 10, 1);
        float ms = time_kernel([&] { k_fill1_v2<10, 10, false><<<n, 32>>>(d.units, n, a1, fo); });
        report("fill1_v2<10,10>", 10, 1, n, d, ms, L);
    }
    if (want("v3")) for (int k : per_sm) {
        if (k > 16 || (only_k && k != only_k)) continue;
        const int n = g_sms * k;
        make_units(n, 10, 1);
        float ms = time_kernel([&] { k_fill1_v3<10, 10, false><<<n, 32>>>(d.units, n, a1, fo, d_offsets); });
        report("fill1_v3<10,10>", 10, 1, n, d, ms, L);
    }
    if (want("v4")) for (int k : per_sm) {
        if (k > 16 || (only_k && k != only_k)) continue;
        const int n = g_sms * k;
        make_units(n, 10, 1);
        const TieArgs tie{64.f * 1.1102230246251565e-16f, 1e-4f};
        float ms = time_kernel([&] { k_fill1_v4<10, 10, false><<<n, 32>>>(d.units, n, a1, fo, d_offsets, tie); });
        report("fill1_v4<10,10>", 10, 1, n, d, ms, L);
    }
#endif
    if (want("c6")) for (int k : per_sm) {
        if (k > 16 || (only_k && k != only_k)) continue;
        const int n = g_sms * k;
        const int ns = (L + 32 * 6 - 1) / (32 * 6);
        make_units(n, 6, ns);
        float ms = ns > 1 ? time_kernel([&] { k_fill1_f32<10, 6, true><<<n, 32>>>(d.units, n, a1, fo); })
                          : time_kernel([&] { k_fill1_f32<10, 6, false><<<n, 32>>>(d.units, n, a1, fo); });
        report("fill1<10,6>", 6, ns, n, d, ms, L);
    }
    if (want("f2")) for (int k : per_sm) {
        if (only_k && k != only_k) continue;
        const int n = g_sms * k;
        make_units(n, 10, 1);
        float ms = time_kernel([&] { k_fill2_f32<10, false><<<n, 32>>>(d.units, n, a2, fo); });
        report("fill2<10>", 10, 1, n, d, ms, L);
    }
#ifdef PROBE_V2
    if (want("f23")) for (int k : per_sm) {
        if (only_k && k != only_k) continue;
        const int n = g_sms * k;
        make_units(n, 10, 1);
        float ms = time_kernel([&] { k_fill2_v3<10, false><<<n, 32>>>(d.units, n, a2, fo, d_offsets); });
        report("fill2_v3<10>", 10, 1, n, d, ms, L);
    }
#endif
    return 0;
}
