// ex2_bw.cu — LAB microbenchmark: MUFU.EX2 throughput per SM (ops / clock), alone and mixed with the FFMA + FADD that
// surround it in the vote kernel's inner loops; 16 or 4 warps per SM, 8 independent chains per thread.
#include <cuda_runtime.h>
#include <cstdio>

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int MODE>
__global__ void __launch_bounds__(512, 1) ex2_kernel(int iters, float c, unsigned long long* cycles, float* sink) {
    float a[8], s[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        a[k] = -0.001f * (threadIdx.x + k);
        s[k] = 0.f;
    }
    __syncthreads();
    const unsigned long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (MODE == 0) {
                a[k] = ex2(a[k]) - 1.0f;            // MUFU + FADD (dependent chain per k, 8 chains)
            } else {
                s[k] += ex2(fmaf(a[k], c, -0.5f));  // FFMA -> MUFU -> FADD, as in the vote kernel; inputs independent
                a[k] += 1e-6f;
            }
        }
    }
    const unsigned long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    float r = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) r += a[k] + s[k];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
static void run(const char* name, int threads) {
    const int sms = 148, iters = 20000;
    unsigned long long* cyc;
    float* sink;
    cudaMalloc(&cyc, sms * sizeof(unsigned long long));
    cudaMalloc(&sink, sms * 512 * sizeof(float));
    ex2_kernel<MODE><<<sms, threads>>>(100, 0.127f, cyc, sink);
    ex2_kernel<MODE><<<sms, threads>>>(iters, 0.127f, cyc, sink);
    cudaDeviceSynchronize();
    unsigned long long h;
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-46s %3d threads: %.2f ex2 / clk / SM\n", name, threads, (double)iters * 8 * threads / (double)h);
    cudaFree(cyc);
    cudaFree(sink);
}

int main() {
    run<0>("ex2 + fadd, dependent chains (8 per thread)", 512);
    run<0>("ex2 + fadd, dependent chains (8 per thread)", 128);
    run<1>("ffma -> ex2 -> fadd (vote inner loop mix)", 512);
    run<1>("ffma -> ex2 -> fadd (vote inner loop mix)", 256);
    run<1>("ffma -> ex2 -> fadd (vote inner loop mix)", 128);
    return 0;
}
