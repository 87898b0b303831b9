// Developer tool: which zero-fill variant streams fastest to HBM (for fill_zero_kernel)?  nvcc -O3 -arch=sm_100a
#include <cstdio>
#include <cuda_runtime.h>
__global__ void fill_plain(float4* p, size_t n4) {
    const size_t stride = (size_t)gridDim.x * blockDim.x; const float4 z = make_float4(0, 0, 0, 0);
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n4; i += 4 * stride) { p[i] = z; p[i + stride] = z; p[i + 2 * stride] = z; p[i + 3 * stride] = z; }
    for (; i < n4; i += stride) p[i] = z;
}
__global__ void fill_cs(float4* p, size_t n4) {
    const size_t stride = (size_t)gridDim.x * blockDim.x; const float4 z = make_float4(0, 0, 0, 0);
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n4; i += 4 * stride) { __stcs(p + i, z); __stcs(p + i + stride, z); __stcs(p + i + 2 * stride, z); __stcs(p + i + 3 * stride, z); }
    for (; i < n4; i += stride) __stcs(p + i, z);
}
__global__ void fill_chunk(float4* p, size_t n4) {   // each CTA owns a contiguous chunk
    const size_t per = (n4 + gridDim.x - 1) / gridDim.x; const size_t b = per * blockIdx.x, e = (b + per < n4) ? b + per : n4;
    const float4 z = make_float4(0, 0, 0, 0);
    for (size_t i = b + threadIdx.x; i < e; i += blockDim.x) p[i] = z;
}
__device__ __forceinline__ void st256(float* q) {     // sm_100: 256-bit global store (STG.E.256)
    asm volatile("st.global.v8.f32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" :: "l"(q), "f"(0.0f) : "memory");
}
__global__ void fill_v8(float* p, size_t n8) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n8; i += 4 * stride) { st256(p + 8 * i); st256(p + 8 * (i + stride)); st256(p + 8 * (i + 2 * stride)); st256(p + 8 * (i + 3 * stride)); }
    for (; i < n8; i += stride) st256(p + 8 * i);
}
__global__ void fill_v8_2(float* p, size_t n8) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + stride < n8; i += 2 * stride) { st256(p + 8 * i); st256(p + 8 * (i + stride)); }
    for (; i < n8; i += stride) st256(p + 8 * i);
}
template <typename F> float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); f(); cudaDeviceSynchronize(); float best = 1e9;
    for (int r = 0; r < 10; r++) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    return best;
}
int main() {
    const size_t bytes = 1152ull << 20; float4* p; cudaMalloc(&p, bytes); const size_t n4 = bytes / 16;
    for (int ctas : {148 * 8, 148 * 32, 148 * 64}) {
        printf("ctas %5d  plain %.1f GB/s  cs %.1f GB/s  chunk %.1f GB/s\n", ctas,
               bytes / timeit([&] { fill_plain<<<ctas, 256>>>(p, n4); }) / 1e6, bytes / timeit([&] { fill_cs<<<ctas, 256>>>(p, n4); }) / 1e6,
               bytes / timeit([&] { fill_chunk<<<ctas, 256>>>(p, n4); }) / 1e6);
    }
    for (int ctas : {148 * 4, 148 * 8, 148 * 16, 148 * 32, 148 * 64})
        for (int thr : {128, 256, 512})
            printf("v8 ctas %5d thr %3d  x4 %.1f GB/s  x2 %.1f GB/s\n", ctas, thr,
                   bytes / timeit([&] { fill_v8<<<ctas, thr>>>((float*)p, n4 / 2); }) / 1e6,
                   bytes / timeit([&] { fill_v8_2<<<ctas, thr>>>((float*)p, n4 / 2); }) / 1e6);
    printf("cudaMemsetAsync %.1f GB/s\n", bytes / timeit([&] { cudaMemsetAsync(p, 0, bytes); }) / 1e6);
    return 0;
}
