// Fused dense losses on [C,H,W] heatmaps, sm_100a.
//
// The reference evaluates l2_loss_gaussian (utils/loss_utils.py:86-100) as ~6 dense ATen passes
// forward (two compares, an OR, sub, pow, a discarded mean, a boolean gather that syncs, mean) and
// ~5 backward.  Here forward is ONE pass (2 loads/element, optionally 1 store) and backward ONE pass
// (2 loads + 1 store); both are pure HBM streams: 128-bit loads, grid = 148 SMs x 8 CTAs.
#include "api_internal.h"

namespace ssb {

constexpr int LOSS_THREADS = 256;

template <int KIND>
__device__ __forceinline__ void loss_elem(float r, float g, float& sum, float& cnt, float& err) {
    const float d = r - g;
    if (KIND == SSB_LOSS_L2_GAUSSIAN) {
        err = d * d;
        if (g > 0.f || r > 0.f) { sum += err; cnt += 1.f; }
    } else if (KIND == SSB_LOSS_L1) {
        err = fabsf(d);
        sum += err; cnt += 1.f;
    } else {
        err = fabsf(d);
        if (g > 0.f || r > 0.f) { sum += err; cnt += 1.f; }
    }
}

template <int KIND>
__global__ void __launch_bounds__(LOSS_THREADS)
loss_fwd_kernel(int64_t n, const float* __restrict__ render, const float* __restrict__ gt,
                double* __restrict__ sums, float* __restrict__ error_out)
{
    // per-thread partials in fp32 over <= a few thousand elements, then fp64 across the grid
    float sum = 0.f, cnt = 0.f;
    const int64_t n4 = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const bool aligned = ((reinterpret_cast<uintptr_t>(render) | reinterpret_cast<uintptr_t>(gt) |
                           reinterpret_cast<uintptr_t>(error_out)) & 15) == 0;
    double dsum = 0.0, dcnt = 0.0;
    if (aligned) {
        const float4* r4 = reinterpret_cast<const float4*>(render);
        const float4* g4 = reinterpret_cast<const float4*>(gt);
        int it = 0;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
            const float4 r = __ldcs(r4 + i), g = __ldcs(g4 + i);
            float4 e;
            loss_elem<KIND>(r.x, g.x, sum, cnt, e.x);
            loss_elem<KIND>(r.y, g.y, sum, cnt, e.y);
            loss_elem<KIND>(r.z, g.z, sum, cnt, e.z);
            loss_elem<KIND>(r.w, g.w, sum, cnt, e.w);
            if (error_out) __stcs(reinterpret_cast<float4*>(error_out) + i, e);
            if (++it == 256) { dsum += sum; dcnt += cnt; sum = 0.f; cnt = 0.f; it = 0; }
        }
        for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
            float e;
            loss_elem<KIND>(render[i], gt[i], sum, cnt, e);
            if (error_out) error_out[i] = e;
        }
    } else {
        int it = 0;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
            float e;
            loss_elem<KIND>(render[i], gt[i], sum, cnt, e);
            if (error_out) error_out[i] = e;
            if (++it == 1024) { dsum += sum; dcnt += cnt; sum = 0.f; cnt = 0.f; it = 0; }
        }
    }
    dsum += sum; dcnt += cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        dsum += __shfl_xor_sync(0xFFFFFFFFu, dsum, o);
        dcnt += __shfl_xor_sync(0xFFFFFFFFu, dcnt, o);
    }
    __shared__ double s_sum[LOSS_THREADS / 32], s_cnt[LOSS_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { s_sum[warp] = dsum; s_cnt[warp] = dcnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, c = 0.0;
#pragma unroll
        for (int w = 0; w < LOSS_THREADS / 32; w++) { a += s_sum[w]; c += s_cnt[w]; }
        atomicAdd(sums, a);
        atomicAdd(sums + 1, c);
    }
}

template <int KIND>
__device__ __forceinline__ float loss_grad_elem(float r, float g, float scale) {
    const float d = r - g;
    if (KIND == SSB_LOSS_L2_GAUSSIAN) return (g > 0.f || r > 0.f) ? 2.f * d * scale : 0.f;
    const float s = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);
    if (KIND == SSB_LOSS_L1) return s * scale;
    return (g > 0.f || r > 0.f) ? s * scale : 0.f;
}

template <int KIND>
__global__ void __launch_bounds__(LOSS_THREADS)
loss_bwd_kernel(int64_t n, const float* __restrict__ render, const float* __restrict__ gt,
                const double* __restrict__ sums, const float* __restrict__ grad_out, float* __restrict__ grad)
{
    const float go = grad_out ? *grad_out : 1.f;
    const double cnt = (KIND == SSB_LOSS_L1) ? (double)n : sums[1];
    const float scale = (float)((double)go / cnt);
    const int64_t n4 = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const bool aligned = ((reinterpret_cast<uintptr_t>(render) | reinterpret_cast<uintptr_t>(gt) |
                           reinterpret_cast<uintptr_t>(grad)) & 15) == 0;
    if (aligned) {
        const float4* r4 = reinterpret_cast<const float4*>(render);
        const float4* g4 = reinterpret_cast<const float4*>(gt);
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
            const float4 r = __ldcs(r4 + i), g = __ldcs(g4 + i);
            float4 o;
            o.x = loss_grad_elem<KIND>(r.x, g.x, scale);
            o.y = loss_grad_elem<KIND>(r.y, g.y, scale);
            o.z = loss_grad_elem<KIND>(r.z, g.z, scale);
            o.w = loss_grad_elem<KIND>(r.w, g.w, scale);
            reinterpret_cast<float4*>(grad)[i] = o;   // default policy: the rasteriser backward re-reads active tiles
        }
        for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
            grad[i] = loss_grad_elem<KIND>(render[i], gt[i], scale);
    } else {
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
            grad[i] = loss_grad_elem<KIND>(render[i], gt[i], scale);
    }
}

// limb_3d_consistency_loss: | |x_a0-x_a1| - |x_b0-x_b1| | + | |x_c0-x_c1| - |x_d0-x_d1| |
struct LimbPairs { int p[8]; };

__device__ __forceinline__ float limb_len(const float* x, int a, int b, float* d) {
    d[0] = x[3 * a] - x[3 * b]; d[1] = x[3 * a + 1] - x[3 * b + 1]; d[2] = x[3 * a + 2] - x[3 * b + 2];
    return sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
}

__global__ void limb_consistency_kernel(int F, int J, const float* __restrict__ xyz, LimbPairs lp,
                                        float* __restrict__ loss, float* __restrict__ grad)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const float* x = xyz + (size_t)f * J * 3;
    float* g = grad ? grad + (size_t)f * J * 3 : nullptr;
    if (g) for (int i = 0; i < J * 3; i++) g[i] = 0.f;
    float total = 0.f;
    for (int k = 0; k < 2; k++) {
        const int a0 = lp.p[4 * k], a1 = lp.p[4 * k + 1], b0 = lp.p[4 * k + 2], b1 = lp.p[4 * k + 3];
        float da[3], db[3];
        const float la = limb_len(x, a0, a1, da), lb = limb_len(x, b0, b1, db);
        const float diff = la - lb;
        total += fabsf(diff);
        if (g) {
            const float s = (diff > 0.f) ? 1.f : ((diff < 0.f) ? -1.f : 0.f);
            for (int c = 0; c < 3; c++) {
                const float ga = (la > 0.f) ? s * da[c] / la : 0.f;
                const float gb = (lb > 0.f) ? -s * db[c] / lb : 0.f;
                g[3 * a0 + c] += ga; g[3 * a1 + c] -= ga;
                g[3 * b0 + c] += gb; g[3 * b1 + c] -= gb;
            }
        }
    }
    if (loss) loss[f] = total;
}

static int loss_grid(int64_t n) {
    int64_t blocks = (n / 4 + LOSS_THREADS - 1) / LOSS_THREADS;
    const int64_t cap = 148 * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

// ------------------------------------------------------------------------------------------ Adam (graph-capturable)
// One optimiser step for the four parameter tensors of ONE frame, exactly as torch.optim.Adam's default foreach path
// (scene/gaussian_model.py:217-218, train.py:215-222): exp_avg.lerp_(g, 1 - b1); exp_avg_sq.mul_(b2).addcmul_(g, g, 1 - b2);
// denom = sqrt(exp_avg_sq) / sqrt(bc2) + eps; p.addcdiv_(exp_avg, denom, -lr / bc1) -- with the step-dependent python-float
// scalars (fp64 on the host, rounded to fp32 where torch hands them to an fp32 tensor op) precomputed for every step in a
// device table and the step index read from (and advanced in) device memory, so the launch can be captured in a CUDA graph
// and replayed: the same arithmetic as phase E of the fused optimiser (optimizer.cu).
// table: [n_steps][5] = -lr_xyz/bc1, -lr_scaling/bc1, -lr_rotation/bc1, -lr_opacity/bc1, sqrt(bc2).
// The xyz gradient is the mean over the V view slots of accumulated_grads (stale / zero slots included, train.py:215-218).
#include "adam_form.h"
__global__ void adam_frame_kernel(int J, int V, float* __restrict__ xyz, float* __restrict__ scaling, float* __restrict__ rotation,
                                  float* __restrict__ opacity, const float* __restrict__ accumulated_grads, const float* __restrict__ g_scaling,
                                  const float* __restrict__ g_rotation, const float* __restrict__ g_opacity, float* __restrict__ exp_avg,
                                  float* __restrict__ exp_avg_sq, const float* __restrict__ table, int n_steps, int* __restrict__ step_counter,
                                  float one_minus_beta1, float beta2, float one_minus_beta2, float eps)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int step = *step_counter;
    if (i < J * 11 && step < n_steps) {
        const float* t = table + (size_t)step * 5;
        float g, neg_step;
        float* param;
        if (i < 3 * J) {
            float a = 0.f;
            for (int v = 0; v < V; v++) a += accumulated_grads[(size_t)v * J * 3 + i];
            g = a / (float)V;
            neg_step = t[0]; param = xyz + i;
        } else if (i < 6 * J) { g = g_scaling[i - 3 * J]; neg_step = t[1]; param = scaling + (i - 3 * J); }
        else if (i < 10 * J) { g = g_rotation[i - 6 * J]; neg_step = t[2]; param = rotation + (i - 6 * J); }
        else { g = g_opacity[i - 10 * J]; neg_step = t[3]; param = opacity + (i - 10 * J); }
        const float m = fmaf(one_minus_beta1, g - exp_avg[i], exp_avg[i]);
        const float vv = SSB_ADAM_SECOND_MOMENT(one_minus_beta2, g, exp_avg_sq[i] * beta2);
        exp_avg[i] = m; exp_avg_sq[i] = vv;
        const float denom = sqrtf(vv) / t[4] + eps;
        *param = fmaf(neg_step, m / denom, *param);
    }
    __syncthreads();                      // every thread has read the step index (one CTA: J * 11 <= 220 threads)
    if (i == 0) *step_counter = step + 1;
}

}  // namespace ssb

using namespace ssb;

extern "C" {

int ssb_loss_forward(int kind, int64_t n, const float* render, const float* gt, double* sums, float* error_out, void* stream_) {
    if (n < 0 || !sums || (n > 0 && (!render || !gt))) return SSB_ERR_INVALID;
    if (n == 0) return SSB_OK;
    cudaStream_t s = (cudaStream_t)stream_;
    const int grid = loss_grid(n);
    switch (kind) {
        case SSB_LOSS_L2_GAUSSIAN: loss_fwd_kernel<SSB_LOSS_L2_GAUSSIAN><<<grid, LOSS_THREADS, 0, s>>>(n, render, gt, sums, error_out); break;
        case SSB_LOSS_L1:          loss_fwd_kernel<SSB_LOSS_L1><<<grid, LOSS_THREADS, 0, s>>>(n, render, gt, sums, error_out); break;
        case SSB_LOSS_L1_GAUSSIAN: loss_fwd_kernel<SSB_LOSS_L1_GAUSSIAN><<<grid, LOSS_THREADS, 0, s>>>(n, render, gt, sums, error_out); break;
        default: return SSB_ERR_INVALID;
    }
    return ssb_set_cuda_error(cudaGetLastError());
}

int ssb_loss_backward(int kind, int64_t n, const float* render, const float* gt, const double* sums,
                      const float* grad_out, float* grad, void* stream_) {
    if (n < 0 || !sums || (n > 0 && (!render || !gt || !grad))) return SSB_ERR_INVALID;
    if (n == 0) return SSB_OK;
    cudaStream_t s = (cudaStream_t)stream_;
    const int grid = loss_grid(n);
    switch (kind) {
        case SSB_LOSS_L2_GAUSSIAN: loss_bwd_kernel<SSB_LOSS_L2_GAUSSIAN><<<grid, LOSS_THREADS, 0, s>>>(n, render, gt, sums, grad_out, grad); break;
        case SSB_LOSS_L1:          loss_bwd_kernel<SSB_LOSS_L1><<<grid, LOSS_THREADS, 0, s>>>(n, render, gt, sums, grad_out, grad); break;
        case SSB_LOSS_L1_GAUSSIAN: loss_bwd_kernel<SSB_LOSS_L1_GAUSSIAN><<<grid, LOSS_THREADS, 0, s>>>(n, render, gt, sums, grad_out, grad); break;
        default: return SSB_ERR_INVALID;
    }
    return ssb_set_cuda_error(cudaGetLastError());
}

int ssb_limb_consistency(int n_frames, int J, const float* xyz, const int* pairs_host, float* loss, float* grad, void* stream_) {
    if (n_frames < 0 || J <= 0 || !pairs_host || (n_frames > 0 && !xyz)) return SSB_ERR_INVALID;
    if (n_frames == 0) return SSB_OK;
    LimbPairs lp;
    // pairs_host: l_arm(a,b) r_arm(a,b) l_leg(a,b) r_leg(a,b)
    for (int i = 0; i < 8; i++) {
        if (pairs_host[i] < 0 || pairs_host[i] >= J) return SSB_ERR_INVALID;
        lp.p[i] = pairs_host[i];
    }
    limb_consistency_kernel<<<(n_frames + 127) / 128, 128, 0, (cudaStream_t)stream_>>>(n_frames, J, xyz, lp, loss, grad);
    return ssb_set_cuda_error(cudaGetLastError());
}

int ssb_adam_frame_step(int J, int V, float* xyz, float* scaling, float* rotation, float* opacity, const float* accumulated_grads,
                        const float* g_scaling, const float* g_rotation, const float* g_opacity, float* exp_avg, float* exp_avg_sq,
                        const float* step_table, int n_steps, int* step_counter, float one_minus_beta1, float beta2, float one_minus_beta2,
                        float eps, void* stream_) {
    if (J <= 0 || J * 11 > 1024 || V <= 0 || n_steps <= 0) return SSB_ERR_INVALID;
    if (!xyz || !scaling || !rotation || !opacity || !accumulated_grads || !g_scaling || !g_rotation || !g_opacity || !exp_avg || !exp_avg_sq ||
        !step_table || !step_counter) return SSB_ERR_INVALID;
    adam_frame_kernel<<<1, ((J * 11 + 31) / 32) * 32, 0, (cudaStream_t)stream_>>>(J, V, xyz, scaling, rotation, opacity, accumulated_grads, g_scaling,
                                                                                 g_rotation, g_opacity, exp_avg, exp_avg_sq, step_table, n_steps,
                                                                                 step_counter, one_minus_beta1, beta2, one_minus_beta2, eps);
    return ssb_set_cuda_error(cudaGetLastError());
}

}  // extern "C"
