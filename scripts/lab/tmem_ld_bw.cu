// tmem_ld_bw.cu — LAB microbenchmark: how fast can an SM read TMEM with tcgen05.ld?
// One CTA per SM, 16 warps (4 per TMEM lane quarter, as the vote kernel's math groups), every warp loads 32 lanes x 16
// fp32 columns per instruction, `iters` times, columns walking over a 128-column accumulator.  Reports bytes / clock / SM.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tmem_ld_bw tmem_ld_bw.cu && ./tmem_ld_bw
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int WARPS, int DEPTH>
__global__ void __launch_bounds__(WARPS * 32, 1) tmem_ld_kernel(int iters, unsigned long long* cycles, uint32_t* sink) {
    __shared__ uint32_t s_tmem;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = s_tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) & 3) * 128;
    uint32_t acc = 0;
    __syncthreads();
    const unsigned long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        uint32_t r[DEPTH][16];
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) {
            const uint32_t a = base + (uint32_t)(((it * DEPTH + d) & 7) * 16);
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(r[d][0]), "=r"(r[d][1]), "=r"(r[d][2]), "=r"(r[d][3]), "=r"(r[d][4]), "=r"(r[d][5]), "=r"(r[d][6]),
                           "=r"(r[d][7]), "=r"(r[d][8]), "=r"(r[d][9]), "=r"(r[d][10]), "=r"(r[d][11]), "=r"(r[d][12]),
                           "=r"(r[d][13]), "=r"(r[d][14]), "=r"(r[d][15])
                         : "r"(a)
                         : "memory");
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int d = 0; d < DEPTH; ++d)
#pragma unroll
            for (int j = 0; j < 16; ++j) acc ^= r[d][j];
    }
    const unsigned long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "r"(512) : "memory");
}

template <int WARPS, int DEPTH>
static void run(const char* name) {
    const int sms = 148, iters = 20000;
    unsigned long long* cyc;
    uint32_t* sink;
    cudaMalloc(&cyc, sms * sizeof(unsigned long long));
    cudaMalloc(&sink, sms * WARPS * 32 * sizeof(uint32_t));
    tmem_ld_kernel<WARPS, DEPTH><<<sms, WARPS * 32>>>(200, cyc, sink);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a);
    tmem_ld_kernel<WARPS, DEPTH><<<sms, WARPS * 32>>>(iters, cyc, sink);
    cudaEventRecord(b);
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    unsigned long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const double bytes = (double)iters * DEPTH * WARPS * 32 * 16 * 4;  // per SM
    printf("%-28s %s  %.1f B/clk/SM  (%.2f TB/s chip-wide at this clock, %.3f ms)\n", name, cudaGetErrorString(e),
           bytes / (double)h[0], bytes * sms / (ms * 1e-3) / 1e12, ms);
    cudaFree(cyc);
    cudaFree(sink);
}

int main() {
    run<16, 1>("16 warps, 1 load in flight");
    run<16, 2>("16 warps, 2 loads in flight");
    run<16, 4>("16 warps, 4 loads in flight");
    run<8, 2>("8 warps, 2 loads in flight");
    run<4, 2>("4 warps, 2 loads in flight");
    run<4, 4>("4 warps, 4 loads in flight");
    return 0;
}
