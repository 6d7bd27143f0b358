// TMEM load / store throughput per SM: 4 warps (one per lane quarter), back-to-back tcgen05.ld / st of 32x32b.x16 / x32
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)
#define R16(v) "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
#define W16(v) "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
template <int MODE>   // 0: ld x16 back to back, wait every 4; 1: st x16; 2: ld + st alternating; 3: ld x16 + wait each (latency)
__global__ void __launch_bounds__(128) k(float *out, int iters, long long *cyc)
{
    __shared__ uint32_t base_s;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&base_s)), "n"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t t = base_s + ((uint32_t)(warp * 32) << 16);
    uint32_t v[16], acc = 0;
    for (int c = 0; c < 16; ++c) v[c] = threadIdx.x + c;
    for (int c = 0; c < 256; c += 16) asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(t + c), W16(v) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t addr = t + ((i * 4 + q) & 15) * 16;
            if (MODE == 0 || MODE == 2 || MODE == 3) {
                uint32_t r[16];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" : R16(r) : "r"(addr));
                if (MODE == 3) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (MODE != 3 && q == 3) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                acc ^= r[q];
            }
            if (MODE == 1 || MODE == 2) asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(addr), W16(v) : "memory");
        }
        if (MODE == 1 || MODE == 2) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    if (acc == 0x12345u) out[0] = acc;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base_s), "n"(256));
}
int main()
{
    long long *d_cyc, h; CK(cudaMalloc(&d_cyc, 8));
    const int iters = 4096;
    for (int ctas : {1, 2}) {
        const char *names[4] = {"ld x16 (wait every 4)", "st x16", "ld + st x16", "ld x16 + wait each"};
        for (int mode = 0; mode < 4; ++mode) {
            if (mode == 0) k<0><<<148 * ctas, 128>>>(nullptr, iters, d_cyc);
            if (mode == 1) k<1><<<148 * ctas, 128>>>(nullptr, iters, d_cyc);
            if (mode == 2) k<2><<<148 * ctas, 128>>>(nullptr, iters, d_cyc);
            if (mode == 3) k<3><<<148 * ctas, 128>>>(nullptr, iters, d_cyc);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(&h, d_cyc, 8, cudaMemcpyDeviceToHost));
            const double instr = (double)iters * 4 * (mode == 2 ? 2 : 1);
            printf("%d CTA/SM  %-24s %7.1f cycles per x16 instruction per warp -> %6.1f B/clk/SM\n", ctas, names[mode], h / instr, 4.0 * ctas * 2048.0 / (h / instr));
        }
    }
    return 0;
}
