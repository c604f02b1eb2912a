// ex2_mix_bw.cu — LAB microbenchmark: does the MUFU.EX2 rate of the vote kernel's inner loop (ffma -> ex2 -> fadd, 16
// scores per step) survive the other MIO traffic of that loop — one LDS.128 per 4 scores (the row statistics) and one
// tcgen05.ld (32 lanes x 16 columns) + tcgen05.wait::ld per 16 scores?  16 warps per SM like the kernel's math groups.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// MODE bit 0: LDS.128 per 4 scores; bit 1: tcgen05.ld.x16 + wait per 16 scores (double-buffered like the kernel)
template <int MODE>
__global__ void __launch_bounds__(512, 1) mix_kernel(int iters, float c, unsigned long long* cycles, float* sink) {
    __shared__ uint32_t s_tmem;
    __shared__ __align__(16) float s_m[128];
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x < 128) s_m[threadIdx.x] = 0.001f * threadIdx.x;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = s_tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) & 3) * 128;
    float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
    uint32_t r[2][16];
#pragma unroll
    for (int j = 0; j < 16; ++j) r[0][j] = r[1][j] = __float_as_uint(-0.01f * j);
    if (MODE & 2)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0][0]), "=r"(r[0][1]), "=r"(r[0][2]), "=r"(r[0][3]), "=r"(r[0][4]), "=r"(r[0][5]), "=r"(r[0][6]),
                       "=r"(r[0][7]), "=r"(r[0][8]), "=r"(r[0][9]), "=r"(r[0][10]), "=r"(r[0][11]), "=r"(r[0][12]), "=r"(r[0][13]),
                       "=r"(r[0][14]), "=r"(r[0][15]) : "r"(base) : "memory");
    __syncthreads();
    const unsigned long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            uint32_t(&v)[16] = r[half];
            if (MODE & 2) {
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                uint32_t(&n)[16] = r[half ^ 1];
                const uint32_t a = base + (uint32_t)(((it * 2 + half + 1) & 7) * 16);
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                             : "=r"(n[0]), "=r"(n[1]), "=r"(n[2]), "=r"(n[3]), "=r"(n[4]), "=r"(n[5]), "=r"(n[6]), "=r"(n[7]),
                               "=r"(n[8]), "=r"(n[9]), "=r"(n[10]), "=r"(n[11]), "=r"(n[12]), "=r"(n[13]), "=r"(n[14]), "=r"(n[15])
                             : "r"(a) : "memory");
            }
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
                float4 mm = make_float4(0.5f, 0.5f, 0.5f, 0.5f);
                if (MODE & 1) {
                    const int o = (it * 32 + half * 16 + j) & 124;
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(mm.x), "=f"(mm.y), "=f"(mm.z), "=f"(mm.w) : "r"(smem_u32(s_m + o)));
                }
                // non-finite TMEM garbage is fine: only the instruction stream matters
                v0 += ex2(fmaf(__uint_as_float(v[j + 0]), c, -mm.x));
                v1 += ex2(fmaf(__uint_as_float(v[j + 1]), c, -mm.y));
                v2 += ex2(fmaf(__uint_as_float(v[j + 2]), c, -mm.z));
                v3 += ex2(fmaf(__uint_as_float(v[j + 3]), c, -mm.w));
            }
        }
    }
    const unsigned long long t1 = clock64();
    if (MODE & 2) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = (v0 + v1) + (v2 + v3) + __uint_as_float(r[0][3] ^ r[1][5]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "r"(512) : "memory");
}

template <int MODE>
static void run(const char* name) {
    const int sms = 148, iters = 20000;
    unsigned long long* cyc;
    float* sink;
    cudaMalloc(&cyc, sms * sizeof(unsigned long long));
    cudaMalloc(&sink, sms * 512 * sizeof(float));
    mix_kernel<MODE><<<sms, 512>>>(100, 0.127f, cyc, sink);
    mix_kernel<MODE><<<sms, 512>>>(iters, 0.127f, cyc, sink);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long h;
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-62s %s: %.2f ex2 / clk / SM\n", name, cudaGetErrorString(e), (double)iters * 32 * 512 / (double)h);
    cudaFree(cyc);
    cudaFree(sink);
}

int main() {
    run<0>("ffma -> ex2 -> fadd, 16 scores per step");
    run<1>("+ one LDS.128 per 4 scores");
    run<2>("+ one tcgen05.ld.x16 + wait::ld per 16 scores");
    run<3>("+ both (the vote kernel's pass-2 loop)");
    return 0;
}
