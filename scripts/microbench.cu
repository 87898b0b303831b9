// Developer tool: measured non-tensor peaks of the chip (fp32 FMA, MUFU ex2, issue rate) for the compute-side fractions
// quoted in profiles/README.md.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a scripts/microbench.cu -o gpurun_out/microbench
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(1024) k(float* out, int iters) {
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
    const float b = 0.999f, c = 1e-3f;
    for (int i = 0; i < iters; i++) {
        if (MODE == 0) {
#pragma unroll
            for (int u = 0; u < 16; u++) {
                a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
                a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
            }
        } else {
#pragma unroll
            for (int u = 0; u < 16; u++) {
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a0)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a1));
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a2)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a3));
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a4)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a5));
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a6)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a7));
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

template <int MODE>
double run(int iters) {
    float* out; cudaMalloc(&out, 148 * 2 * 1024 * sizeof(float));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 2, 1024>>>(out, iters); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0); k<MODE><<<148 * 2, 1024>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    cudaFree(out);
    const double ops = 148.0 * 2 * 1024 * (double)iters * 16 * 8;
    return ops / (best * 1e-3);
}

int main() {
    const double ffma = run<0>(4096), ex2 = run<1>(2048);
    printf("{\"fp32_fma_tflops\": %.2f, \"ffma_per_s\": %.4e, \"mufu_ex2_per_s\": %.4e, \"warp_instr_issue_peak_per_s_nominal\": %.4e}\n",
           2 * ffma / 1e12, ffma, ex2, 148.0 * 4 * 1.965e9);
    return 0;
}
