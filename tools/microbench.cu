// Issue/pipe throughput probes for the instructions the pair kernels are made of (sm_100a).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu ; prints warp-instr/clk/SM.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
#define UN 8
template <int MODE> __global__ void probe(float *out, int iters, float seed)
{
    float x[UN]; float2 p[UN];
    for (int q = 0; q < UN; ++q) { x[q] = seed + threadIdx.x * 1e-3f + q; p[q] = make_float2(x[q], x[q] + 0.5f); }
    const float a = 0.999f, b = 1e-3f; const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    unsigned w = 0;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int q = 0; q < UN; ++q) {
                if (MODE == 0) x[q] = __fmaf_rn(x[q], a, b);
                if (MODE == 1) p[q] = __ffma2_rn(p[q], a2, b2);
                if (MODE == 2) x[q] = fmaxf(fmaxf(x[q], a), b + x[(q + 1) % UN]);           // FMNMX3 (+FADD)
                if (MODE == 3) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x[q])); x[q] = y; }
                if (MODE == 4) { unsigned m; asm volatile("set.neu.u32.f32 %0, %1, %2;" : "=r"(m) : "f"(x[q]), "f"(x[(q + 1) % UN])); w = (w | (m & (3u << (2 * q)))); x[q] += 1.0f; }
                if (MODE == 5) { x[q] = __fmaf_rn(x[q], a, b); p[q] = __ffma2_rn(p[q], a2, b2); }          // mix 1:1
                if (MODE == 6) { x[q] = fmaxf(x[q], x[(q + 1) % UN] + b); }                               // FADD + FMNMX
                if (MODE == 7) { p[q] = __fadd2_rn(p[q], b2); }
                if (MODE == 8) { x[q] = __fmaf_rn(x[q], a, b); x[q] = fmaxf(x[q], b); }                  // FFMA + FMNMX (fma + alu pipes)
                if (MODE == 9) { p[q] = __ffma2_rn(p[q], a2, b2); x[q] = fmaxf(x[q], b + p[q].x); }
                if (MODE == 10) { p[q] = __ffma2_rn(p[q], a2, b2); w = (w ^ __float_as_uint(x[q])) + (unsigned)q; }            // FFMA2 + LOP3/IADD (alu)
                if (MODE == 13) { p[q] = __ffma2_rn(make_float2(x[q], x[q]), a2, p[q]); }                                      // FFMA2 scalar-broadcast operand
                if (MODE == 16) { x[q] = __fmaf_rn(x[q], a, b); w = (w ^ __float_as_uint(x[(q + 3) % UN])) + (unsigned)q; }     // FFMA + alu
                if (MODE == 17) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[q])); p[q] = __ffma2_rn(p[q], a2, b2); p[(q + 1) % UN] = __ffma2_rn(p[(q + 1) % UN], a2, b2); p[(q + 2) % UN] = __ffma2_rn(p[(q + 2) % UN], a2, b2); }
                if (MODE == 18) { w = __funnelshift_l(__float_as_uint(x[q]), w, 1); x[q] += b; }                              // SHF + FADD
                if (MODE == 19) { x[q] = __fmaf_rn(x[q], a, b); x[q] = __fmaf_rn(x[q], a, b); x[q] = __fmaf_rn(x[q], a, b); x[q] = fmaxf(x[q], b); w = __funnelshift_l(__float_as_uint(x[q]), w, 1); }  // 3 FFMA + FMNMX + SHF
            }
        }
    }
    float s = 0; for (int q = 0; q < UN; ++q) s += x[q] + p[q].x + p[q].y;
    if (s == 1234.5f) out[0] = s + w;
}
// order pinned with asm volatile: 8 packed FMAs and 8 scalar adds per inner body, grouped (4+4 runs) or alternating
template <int GROUP> __global__ void probe_mix(float *out, int iters, float seed)
{
    float x[8]; unsigned long long p[8];
    for (int q = 0; q < 8; ++q) { x[q] = seed + threadIdx.x * 1e-3f + q; float2 t = make_float2(x[q], x[q] + 0.5f); p[q] = *reinterpret_cast<unsigned long long *>(&t); }
    float2 a2f = make_float2(0.999f, 0.999f), b2f = make_float2(1e-3f, 1e-3f);
    unsigned long long a2 = *reinterpret_cast<unsigned long long *>(&a2f), b2 = *reinterpret_cast<unsigned long long *>(&b2f);
    const float b = 1e-3f;
#define PK(q) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[q]) : "l"(a2), "l"(b2));
#define SC(q) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x[q]) : "f"(b));
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            if (GROUP == 1) { PK(0) SC(0) PK(1) SC(1) PK(2) SC(2) PK(3) SC(3) PK(4) SC(4) PK(5) SC(5) PK(6) SC(6) PK(7) SC(7) }
            if (GROUP == 4) { PK(0) PK(1) PK(2) PK(3) SC(0) SC(1) SC(2) SC(3) PK(4) PK(5) PK(6) PK(7) SC(4) SC(5) SC(6) SC(7) }
            if (GROUP == 8) { PK(0) PK(1) PK(2) PK(3) PK(4) PK(5) PK(6) PK(7) SC(0) SC(1) SC(2) SC(3) SC(4) SC(5) SC(6) SC(7) }
        }
    }
    float s = 0; for (int q = 0; q < 8; ++q) { float2 t = *reinterpret_cast<float2 *>(&p[q]); s += x[q] + t.x + t.y; }
    if (s == 1234.5f) out[0] = s;
}
template <int GROUP> void run_mix(const char *name)
{
    float *d; cudaMalloc(&d, 64);
    int dev; cudaGetDevice(&dev); cudaDeviceProp pr; cudaGetDeviceProperties(&pr, dev);
    int grid = pr.multiProcessorCount * 8, block = 256;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe_mix<GROUP><<<grid, block>>>(d, 16, 1.0f);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0); probe_mix<GROUP><<<grid, block>>>(d, ITERS, 1.0f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double warp_instr = (double)grid * (block / 32) * (double)ITERS * 8 * 16;
    printf("%-28s %8.3f ms  %7.3f warp-instr/clk/SM (at 1.9 GHz nominal)\n", name, best, warp_instr / (best * 1e-3) / 1.9e9 / pr.multiProcessorCount);
    cudaFree(d);
}
template <int MODE> void run(const char *name, double instr_per_inner)
{
    float *d; cudaMalloc(&d, 64);
    int dev; cudaGetDevice(&dev); cudaDeviceProp pr; cudaGetDeviceProperties(&pr, dev);
    int grid = pr.multiProcessorCount * 8, block = 256;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<MODE><<<grid, block>>>(d, 16, 1.0f);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0); probe<MODE><<<grid, block>>>(d, ITERS, 1.0f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double warp_instr = (double)grid * (block / 32) * (double)ITERS * 8 * UN * instr_per_inner;
    double clk = 1.9e9;   // nominal; ratios between rows are what matters
    printf("%-28s %8.3f ms  %7.3f warp-instr/clk/SM (at 1.9 GHz nominal)\n", name, best, warp_instr / (best * 1e-3) / clk / pr.multiProcessorCount);
    cudaFree(d);
}
int main()
{
    run<0>("FFMA", 1); run<1>("FFMA2", 1); run<2>("FMNMX3+FADD", 2); run<3>("MUFU.EX2", 1); run<4>("FSET+LOP3+FADD", 3);
    run<5>("FFMA+FFMA2", 2); run<6>("FADD+FMNMX", 2); run<7>("FADD2", 1); run<8>("FFMA+FMNMX", 2); run<9>("FFMA2+FADD+FMNMX", 3);
    run<10>("FFMA2+LOP3+IADD", 3); run<13>("FFMA2 scalar-bcast", 1); run<16>("FFMA+LOP3+IADD", 3); run<17>("MUFU+3 FFMA2", 4); run<18>("SHF+FADD", 2);
    run<19>("3 FFMA+FMNMX+SHF", 5);
    run_mix<1>("FFMA2/FADD alternating"); run_mix<4>("FFMA2/FADD groups of 4"); run_mix<8>("FFMA2/FADD groups of 8");
    return 0;
}
