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
        } else if (MODE == 2) {
            // packed fp32 (fma.rn.f32x2 -> FFMA2): 8 independent chains of 64-bit pairs; counted as ONE op per instruction here
            unsigned long long p0, p1, p2, p3, p4, p5, p6, p7, bb, cc;
            asm("mov.b64 %0, {%1, %2};" : "=l"(p0) : "f"(a0), "f"(a1)); asm("mov.b64 %0, {%1, %2};" : "=l"(p1) : "f"(a2), "f"(a3));
            asm("mov.b64 %0, {%1, %2};" : "=l"(p2) : "f"(a4), "f"(a5)); asm("mov.b64 %0, {%1, %2};" : "=l"(p3) : "f"(a6), "f"(a7));
            asm("mov.b64 %0, {%1, %2};" : "=l"(p4) : "f"(a1), "f"(a0)); asm("mov.b64 %0, {%1, %2};" : "=l"(p5) : "f"(a3), "f"(a2));
            asm("mov.b64 %0, {%1, %2};" : "=l"(p6) : "f"(a5), "f"(a4)); asm("mov.b64 %0, {%1, %2};" : "=l"(p7) : "f"(a7), "f"(a6));
            asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b)); asm("mov.b64 %0, {%1, %1};" : "=l"(cc) : "f"(c));
#pragma unroll
            for (int u = 0; u < 16; u++) {
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(bb), "l"(cc)); asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p1) : "l"(bb), "l"(cc));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p2) : "l"(bb), "l"(cc)); asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p3) : "l"(bb), "l"(cc));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p4) : "l"(bb), "l"(cc)); asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p5) : "l"(bb), "l"(cc));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p6) : "l"(bb), "l"(cc)); asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p7) : "l"(bb), "l"(cc));
            }
            float x, y;
            asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(p0)); a0 = x + y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(p1)); a1 = x + y;
            asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(p2)); a2 = x + y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(p3)); a3 = x + y;
            asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(p4)); a4 = x + y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(p5)); a5 = x + y;
            asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(p6)); a6 = x + y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(p7)); a7 = x + y;
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
    const double ffma = run<0>(4096), ex2 = run<1>(2048), ffma2 = run<2>(4096);
    printf("{\"fp32_fma_tflops\": %.2f, \"ffma_per_s\": %.4e, \"mufu_ex2_per_s\": %.4e, \"warp_instr_issue_peak_per_s_nominal\": %.4e, "
           "\"ffma2_instr_per_s\": %.4e, \"ffma2_fp32_fma_tflops\": %.2f, \"ffma2_instr_rate_vs_ffma\": %.3f}\n",
           2 * ffma / 1e12, ffma, ex2, 148.0 * 4 * 1.965e9, ffma2, 4 * ffma2 / 1e12, ffma2 / ffma);
    return 0;
}
