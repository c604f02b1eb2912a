// copyctl.cu — LAB ONLY (not part of the product library): two controls that move exactly the address set of
// streaming_llm_compress at BASELINE c2 — per (batch, head) unit rows [0, sink) and [S - tail, S) of a [B,H,S,D]
// tensor into a dense [B,H,sink+tail,D] tensor — so that the product kernel's 0.915 of the copy peak can be compared
// with (a) a plain LDG.128/STG.128 kernel and (b) cudaMemcpy2DAsync over the same bytes.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -shared -o libcopyctl.so copyctl.cu
#include <cuda_runtime.h>
#include <stdint.h>

__global__ void __launch_bounds__(256) copyctl_ldg_kernel(const char* __restrict__ in, char* __restrict__ out,
                                                          int64_t unit_stride, int row_bytes, int sink, int tail, int S) {
    const int unit = blockIdx.x;
    const char* src_unit = in + (int64_t)unit * unit_stride;
    char* dst_unit = out + (int64_t)unit * (int64_t)(sink + tail) * row_bytes;
    const int64_t sink_chunks = (int64_t)sink * row_bytes / 16, tail_chunks = (int64_t)tail * row_bytes / 16;
    const char* tail_src = src_unit + (int64_t)(S - tail) * row_bytes;
    const int64_t total = sink_chunks + tail_chunks;
    // this CTA's slice of the unit (gridDim.y CTAs per unit), 4 loads in flight per thread
    const int64_t per = (total + gridDim.y - 1) / gridDim.y;
    const int64_t lo = (int64_t)blockIdx.y * per, hi = lo + per < total ? lo + per : total;
    for (int64_t c0 = lo + threadIdx.x; c0 < hi; c0 += 4 * blockDim.x) {
        int4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t c = c0 + (int64_t)u * blockDim.x;
            if (c < hi) v[u] = *reinterpret_cast<const int4*>(c < sink_chunks ? src_unit + c * 16 : tail_src + (c - sink_chunks) * 16);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t c = c0 + (int64_t)u * blockDim.x;
            if (c < hi) *reinterpret_cast<int4*>(dst_unit + c * 16) = v[u];
        }
    }
}

extern "C" int copyctl_ldg(const void* in, void* out, int units, int64_t unit_stride, int row_bytes, int sink, int tail,
                           int S, int ctas_per_unit, void* stream) {
    dim3 grid((unsigned)units, (unsigned)ctas_per_unit, 1);
    copyctl_ldg_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const char*)in, (char*)out, unit_stride, row_bytes, sink,
                                                              tail, S);
    return (int)cudaGetLastError();
}

extern "C" int copyctl_memcpy2d(const void* in, void* out, int units, int64_t unit_stride, int row_bytes, int sink,
                                int tail, int S, void* stream) {
    const size_t dpitch = (size_t)(sink + tail) * row_bytes;
    cudaError_t e = cudaSuccess;
    if (sink > 0)
        e = cudaMemcpy2DAsync(out, dpitch, in, (size_t)unit_stride, (size_t)sink * row_bytes, (size_t)units,
                              cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
    if (e == cudaSuccess && tail > 0)
        e = cudaMemcpy2DAsync((char*)out + (size_t)sink * row_bytes, dpitch, (const char*)in + (size_t)(S - tail) * row_bytes,
                              (size_t)unit_stride, (size_t)tail * row_bytes, (size_t)units, cudaMemcpyDeviceToDevice,
                              (cudaStream_t)stream);
    return (int)e;
}
