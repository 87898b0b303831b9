// Fused differentiable SSIM (11x11 Gaussian window, sigma 1.5, zero padding), sm_100a.
//
// Same contract as the reference's fused-ssim extension (submodules/fused-ssim/ssim.cu:187-444):
// forward writes the SSIM map and, when training, the three partial-derivative maps
// dm/dmu1, dm/dsigma1^2, dm/dsigma12; backward returns dL/dimg1 as three Gaussian convolutions of
// dL/dmap * dm/d{...}.  Design differences: one CTA per (32x32 tile, image plane) instead of a serial
// loop over channels (B*CH x more CTAs in flight), all five moments convolved in one horizontal and
// one vertical pass out of a single staging of the two 42x42 halo tiles (the reference re-stages
// and re-synchronises per moment: 5 x (flush, conv-x, conv-y) with 20 block syncs; here 3).
#include "api_internal.h"

namespace ssb {

constexpr int SS_T = 32;            // output tile edge
constexpr int SS_R = 5;             // window radius
constexpr int SS_H = SS_T + 2 * SS_R;   // halo tile edge (42)
constexpr int SS_THREADS = 256;

// 11-tap normalised Gaussian, sigma = 1.5: the literal constants of the reference (ssim.cu:9-19)
__constant__ float c_gauss[11] = {
    0.001028380123898387f, 0.0075987582094967365f, 0.036000773310661316f, 0.10936068743467331f,
    0.21300552785396576f, 0.26601171493530273f, 0.21300552785396576f, 0.10936068743467331f,
    0.036000773310661316f, 0.0075987582094967365f, 0.001028380123898387f};

__device__ __forceinline__ float load_zero_pad(const float* __restrict__ plane, int y, int x, int H, int W) {
    return (x >= 0 && y >= 0 && x < W && y < H) ? __ldg(plane + (size_t)y * W + x) : 0.0f;
}

__global__ void __launch_bounds__(SS_THREADS)
ssim_fwd_kernel(int H, int W, float C1, float C2, const float* __restrict__ img1, const float* __restrict__ img2,
                float* __restrict__ ssim_map, float* __restrict__ dm_dmu1, float* __restrict__ dm_dsigma1_sq,
                float* __restrict__ dm_dsigma12)
{
    __shared__ float s1[SS_H][SS_H + 1], s2[SS_H][SS_H + 1];
    __shared__ float h[5][SS_H][SS_T + 1];     // horizontal pass of x, y, x^2, y^2, xy
    const size_t plane = (size_t)blockIdx.z * H * W;
    const float* p1 = img1 + plane;
    const float* p2 = img2 + plane;
    const int x0 = blockIdx.x * SS_T, y0 = blockIdx.y * SS_T;
    const int tid = threadIdx.x;
    for (int i = tid; i < SS_H * SS_H; i += SS_THREADS) {
        const int ly = i / SS_H, lx = i - ly * SS_H;
        s1[ly][lx] = load_zero_pad(p1, y0 + ly - SS_R, x0 + lx - SS_R, H, W);
        s2[ly][lx] = load_zero_pad(p2, y0 + ly - SS_R, x0 + lx - SS_R, H, W);
    }
    __syncthreads();
    for (int i = tid; i < SS_H * SS_T; i += SS_THREADS) {
        const int ly = i / SS_T, lx = i - ly * SS_T;
        float a = 0.f, b = 0.f, aa = 0.f, bb = 0.f, ab = 0.f;
#pragma unroll
        for (int k = 0; k < 11; k++) {
            const float g = c_gauss[k], u = s1[ly][lx + k], v = s2[ly][lx + k];
            a += g * u; b += g * v; aa += g * (u * u); bb += g * (v * v); ab += g * (u * v);
        }
        h[0][ly][lx] = a; h[1][ly][lx] = b; h[2][ly][lx] = aa; h[3][ly][lx] = bb; h[4][ly][lx] = ab;
    }
    __syncthreads();
    for (int i = tid; i < SS_T * SS_T; i += SS_THREADS) {
        const int ly = i / SS_T, lx = i - ly * SS_T;
        const int px = x0 + lx, py = y0 + ly;
        if (px >= W || py >= H) continue;
        float mu1 = 0.f, mu2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
        for (int k = 0; k < 11; k++) {
            const float g = c_gauss[k];
            mu1 += g * h[0][ly + k][lx]; mu2 += g * h[1][ly + k][lx];
            e11 += g * h[2][ly + k][lx]; e22 += g * h[3][ly + k][lx]; e12 += g * h[4][ly + k][lx];
        }
        const float sigma1_sq = e11 - mu1 * mu1, sigma2_sq = e22 - mu2 * mu2, sigma12 = e12 - mu1 * mu2;
        const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu1_mu2 = mu1 * mu2;
        const float Cc = 2.0f * mu1_mu2 + C1, D = 2.0f * sigma12 + C2;
        const float A = mu1_sq + mu2_sq + C1, B = sigma1_sq + sigma2_sq + C2;
        const size_t o = plane + (size_t)py * W + px;
        ssim_map[o] = (Cc * D) / (A * B);
        if (dm_dmu1) {
            dm_dmu1[o] = (mu2 * 2.0f * D) / (A * B) - (mu2 * 2.0f * Cc) / (A * B) - (mu1 * 2.0f * Cc * D) / (A * A * B) + (mu1 * 2.0f * Cc * D) / (A * B * B);
            dm_dsigma1_sq[o] = (-Cc * D) / (A * B * B);
            dm_dsigma12[o] = (2.0f * Cc) / (A * B);
        }
    }
}

__global__ void __launch_bounds__(SS_THREADS)
ssim_bwd_kernel(int H, int W, const float* __restrict__ img1, const float* __restrict__ img2,
                const float* __restrict__ dL_dmap, const float* __restrict__ dm_dmu1,
                const float* __restrict__ dm_dsigma1_sq, const float* __restrict__ dm_dsigma12,
                float* __restrict__ dL_dimg1)
{
    __shared__ float s[3][SS_H][SS_H + 1];      // dL_dmap * dm_d{mu1, sigma1_sq, sigma12}
    __shared__ float h[3][SS_H][SS_T + 1];
    const size_t plane = (size_t)blockIdx.z * H * W;
    const int x0 = blockIdx.x * SS_T, y0 = blockIdx.y * SS_T;
    const int tid = threadIdx.x;
    for (int i = tid; i < SS_H * SS_H; i += SS_THREADS) {
        const int ly = i / SS_H, lx = i - ly * SS_H;
        const int y = y0 + ly - SS_R, x = x0 + lx - SS_R;
        const float dl = load_zero_pad(dL_dmap + plane, y, x, H, W);
        s[0][ly][lx] = dl * load_zero_pad(dm_dmu1 + plane, y, x, H, W);
        s[1][ly][lx] = dl * load_zero_pad(dm_dsigma1_sq + plane, y, x, H, W);
        s[2][ly][lx] = dl * load_zero_pad(dm_dsigma12 + plane, y, x, H, W);
    }
    __syncthreads();
    for (int i = tid; i < SS_H * SS_T; i += SS_THREADS) {
        const int ly = i / SS_T, lx = i - ly * SS_T;
        float a = 0.f, b = 0.f, c = 0.f;
#pragma unroll
        for (int k = 0; k < 11; k++) {
            const float g = c_gauss[k];
            a += g * s[0][ly][lx + k]; b += g * s[1][ly][lx + k]; c += g * s[2][ly][lx + k];
        }
        h[0][ly][lx] = a; h[1][ly][lx] = b; h[2][ly][lx] = c;
    }
    __syncthreads();
    for (int i = tid; i < SS_T * SS_T; i += SS_THREADS) {
        const int ly = i / SS_T, lx = i - ly * SS_T;
        const int px = x0 + lx, py = y0 + ly;
        if (px >= W || py >= H) continue;
        float a = 0.f, b = 0.f, c = 0.f;
#pragma unroll
        for (int k = 0; k < 11; k++) {
            const float g = c_gauss[k];
            a += g * h[0][ly + k][lx]; b += g * h[1][ly + k][lx]; c += g * h[2][ly + k][lx];
        }
        const size_t o = plane + (size_t)py * W + px;
        const float pix1 = __ldg(img1 + o), pix2 = __ldg(img2 + o);
        dL_dimg1[o] = a + pix1 * 2.0f * b + pix2 * c;
    }
}

}  // namespace ssb

using namespace ssb;

extern "C" {

int ssb_fused_ssim_forward(int B, int CH, int H, int W, float C1, float C2, const float* img1, const float* img2,
                           float* ssim_map, float* dm_dmu1, float* dm_dsigma1_sq, float* dm_dsigma12, void* stream_) {
    if (B < 0 || CH < 0 || H < 0 || W < 0) return SSB_ERR_INVALID;
    if ((size_t)B * CH * H * W == 0) return SSB_OK;
    if (!img1 || !img2 || !ssim_map) return SSB_ERR_INVALID;
    if ((dm_dmu1 != nullptr) != (dm_dsigma1_sq != nullptr) || (dm_dmu1 != nullptr) != (dm_dsigma12 != nullptr)) return SSB_ERR_INVALID;
    if ((long long)B * CH > 65535) return SSB_ERR_CAPACITY;
    const dim3 grid((W + SS_T - 1) / SS_T, (H + SS_T - 1) / SS_T, B * CH);
    ssim_fwd_kernel<<<grid, SS_THREADS, 0, (cudaStream_t)stream_>>>(H, W, C1, C2, img1, img2, ssim_map, dm_dmu1, dm_dsigma1_sq, dm_dsigma12);
    return ssb_set_cuda_error(cudaGetLastError());
}

int ssb_fused_ssim_backward(int B, int CH, int H, int W, float C1, float C2, const float* img1, const float* img2,
                            const float* dL_dmap, const float* dm_dmu1, const float* dm_dsigma1_sq,
                            const float* dm_dsigma12, float* dL_dimg1, void* stream_) {
    (void)C1; (void)C2;
    if (B < 0 || CH < 0 || H < 0 || W < 0) return SSB_ERR_INVALID;
    if ((size_t)B * CH * H * W == 0) return SSB_OK;
    if (!img1 || !img2 || !dL_dmap || !dm_dmu1 || !dm_dsigma1_sq || !dm_dsigma12 || !dL_dimg1) return SSB_ERR_INVALID;
    if ((long long)B * CH > 65535) return SSB_ERR_CAPACITY;
    const dim3 grid((W + SS_T - 1) / SS_T, (H + SS_T - 1) / SS_T, B * CH);
    ssim_bwd_kernel<<<grid, SS_THREADS, 0, (cudaStream_t)stream_>>>(H, W, img1, img2, dL_dmap, dm_dmu1, dm_dsigma1_sq, dm_dsigma12, dL_dimg1);
    return ssb_set_cuda_error(cudaGetLastError());
}

}  // extern "C"
