// FP64 pipe probes for the float64 kernels (parity-mode fills, affine DTW fill, node kernels) on sm_100a:
// throughput with 8 independent chains per thread and 8 warps per SM sub-partition, and the dependent-issue latency of one chain in
// one warp.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_fp64 microbench_fp64.cu
#include <cstdio>
#include <cuda_runtime.h>
#define UN 8
template <int MODE> __global__ void probe(double *out, int iters, double seed)
{
    double x[UN];
    for (int q = 0; q < UN; ++q) x[q] = seed + threadIdx.x * 1e-3 + q;
    const double a = 0.999, b = 1e-3;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int q = 0; q < UN; ++q) {
                if (MODE == 0) x[q] = __dadd_rn(x[q], b);                                         // DADD
                if (MODE == 1) x[q] = __fma_rn(x[q], a, b);                                       // DFMA
                if (MODE == 2) x[q] = fmax(x[q], x[(q + 1) % UN]);                                // DSETP + select (or DMNMX)
                if (MODE == 3) { const double t = __dadd_rn(x[q], b); x[q] = t > x[(q + 1) % UN] ? t : x[(q + 1) % UN]; }   // DADD + DSETP + SEL
                if (MODE == 4) {                                                                   // integer compare of the bit patterns (non-negative values)
                    const long long u = __double_as_longlong(x[q]), v = __double_as_longlong(x[(q + 1) % UN]);
                    x[q] = __longlong_as_double(u > v ? u : v + 1);
                }
            }
        }
    }
    double s = 0; for (int q = 0; q < UN; ++q) s += x[q];
    if (s == 1234.5) out[0] = s;
}
template <int MODE> __global__ void chain(double *out, int iters, double seed)                    // one dependent chain, one warp
{
    double x = seed + threadIdx.x * 1e-3, y = seed * 0.5;
    const double b = 1e-3;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 64; ++r) {
            if (MODE == 0) x = __dadd_rn(x, b);
            if (MODE == 2) { x = x > y ? x : y; y = __longlong_as_double(__double_as_longlong(y) + 1); }     // DSETP + SEL on the chain (+ IADD off it)
            if (MODE == 3) { const double t = __dadd_rn(x, b); x = t > y ? t : y; }
        }
    }
    if (x == 1234.5) out[0] = x + y;
}
template <int MODE> void run(const char *name, double instr)
{
    double *d; cudaMalloc(&d, 64);
    int dev; cudaGetDevice(&dev); cudaDeviceProp pr; cudaGetDeviceProperties(&pr, dev);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int warps_per_sm : {1, 4, 32}) {
        int grid = pr.multiProcessorCount, block = 32 * warps_per_sm, iters = 512;
        probe<MODE><<<grid, block>>>(d, 8, 1.0);
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0); probe<MODE><<<grid, block>>>(d, iters, 1.0); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        double wi = (double)grid * warps_per_sm * iters * 8.0 * UN * instr;
        printf("%-34s %2d warps/SM  %8.3f ms  %7.4f warp-instr/clk/SM\n", name, warps_per_sm, best, wi / (best * 1e-3) / (clk * 1e3) / pr.multiProcessorCount);
    }
    cudaFree(d);
}
template <int MODE> void run_chain(const char *name, double instr)
{
    double *d; cudaMalloc(&d, 64);
    int dev; cudaGetDevice(&dev);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    chain<MODE><<<1, 32>>>(d, 8, 1.0);
    float best = 1e30f; const int iters = 4096;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0); chain<MODE><<<1, 32>>>(d, iters, 1.0); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    printf("%-34s dependent chain  %8.3f ms  %7.2f clk per step (%g instr on the chain)\n", name, best, best * 1e-3 * clk * 1e3 / (iters * 64.0), instr);
    cudaFree(d);
}
int main()
{
    run<0>("DADD", 1); run<1>("DFMA", 1); run<2>("fmax(double) = DSETP+sel / DMNMX", 1); run<3>("DADD + compare + select", 1); run<4>("int64 compare + select", 1);
    run_chain<0>("DADD", 1); run_chain<2>("compare + select", 1); run_chain<3>("DADD + compare + select", 2);
    return 0;
}
