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
            }
        }
    }
    float s = 0; for (int q = 0; q < UN; ++q) s += x[q] + p[q].x + p[q].y;
    if (s == 1234.5f) out[0] = s + w;
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
    return 0;
}
