// Fused differentiable SSIM (11x11 Gaussian window, sigma 1.5, zero padding), sm_100a.
//
// Same contract as the reference's fused-ssim extension (submodules/fused-ssim/ssim.cu:187-444):
// forward writes the SSIM map and, when training, the three partial-derivative maps
// dm/dmu1, dm/dsigma1^2, dm/dsigma12; backward returns dL/dimg1 as three Gaussian convolutions of
// dL/dmap * dm/d{...}.
//
// Design (round 2).  The reference (and round 1 here) stage a 42x42 halo tile in shared memory and run both 1-D passes out of
// it: ~80 scalar shared-memory loads per output pixel, which bounds the kernel at ~15-20 % of the HBM roof.  Here every WARP
// owns a 32-column strip and marches down the rows (no block-level barrier at all):
//   * horizontal pass: the row's 42 input pixels of both images go through a per-warp (u,v)-interleaved row buffer, so one
//     LDS.64 per tap delivers the (u,v) pair; the five moments are accumulated with PACKED fp32 math (FFMA2 / FMUL2,
//     `fma.rn.f32x2`, new on sm_100) as the pairs (mu1, mu2) and (E[x^2] + E[y^2], E[xy]) -- SSIM needs the two variances only as
//     their sum; taps k and 10 - k share a weight, so their squares / products are summed first;
//   * vertical pass: the last 11 horizontal results live in a REGISTER ring (the row loop is unrolled by 11 so every ring
//     index is static): 2 FFMA2 per tap, no shared memory, and a finished output row every input row;
//   * input rows arrive by cp.async (4-byte LDGSTS, zero-filling the padding) into a ring of SS_NB row buffers per warp, SS_NB - 1
//     rows in flight; copy addresses come from clamped coordinates with 32-bit index math (one IMAD.WIDE per copy);
//   * the backward kernel gives every lane TWO adjacent columns of a 64-column strip: the 12 ring values per stream serve both
//     and arrive as LDS.128 / LDS.64 (the one-column form was bound by the shared-memory pipe);
//   * `fused_ssim()` only ever consumes map.mean(): the MEAN variants reduce the map in the kernel (per-warp partial sums,
//     summed in a fixed order in fp64 by a tiny second kernel) and take dL/dmap as the scalar it is, so neither the map nor
//     dL/dmap ever touches HBM (train step: 495 MB instead of 675 MB at 5x1x1500x1500).
// Per-pixel epilogue uses two IEEE reciprocals (1/A, 1/B) instead of the reference's seven divisions; results agree with
// the reference's kernels to ~1e-6 relative (tests/test_reference_python.py, tests/test_gpu_losses.py).
#include "api_internal.h"
#include "f32x2.cuh"

namespace ssb {

constexpr int SS_R = 5;              // window radius
#ifndef SS_WARPS_PER_CTA
#define SS_WARPS_PER_CTA 8
#endif
#ifndef SS_ROW_BUFFERS
#define SS_ROW_BUFFERS 8
#endif
#ifndef SS_WARPS_PER_CTA_BWD
#define SS_WARPS_PER_CTA_BWD SS_WARPS_PER_CTA
#endif
#ifndef SS_FWD_RESIDENT_WARPS
#define SS_FWD_RESIDENT_WARPS 16     // same for the forward launch
#endif
#ifndef SS_BWD_RESIDENT_WARPS
#define SS_BWD_RESIDENT_WARPS 16     // warps per SM the backward launch keeps resident (CTAs per SM x warps per CTA): sizes the strips
#endif
constexpr int SS_WARPS = SS_WARPS_PER_CTA;   // warps (= 32-column strips) per CTA, forward
constexpr int SS_WARPS_B = SS_WARPS_PER_CTA_BWD;   // warps (= 64-column strips) per CTA, backward
constexpr int SS_ROWS_MAX = 96;      // output rows per warp strip: chosen per launch (ss_rows) so that the strips fill whole waves
constexpr int SS_BUFW = 48;          // row buffer width (>= 32 + 2*SS_R)
constexpr int SS_NB = SS_ROW_BUFFERS; // row buffers per warp (power of two): SS_NB - 1 rows of cp.async in flight

// 11-tap normalised Gaussian, sigma = 1.5: the literal constants of the reference (ssim.cu:9-19); symmetric
#define SS_G0 0.001028380123898387f
#define SS_G1 0.0075987582094967365f
#define SS_G2 0.036000773310661316f
#define SS_G3 0.10936068743467331f
#define SS_G4 0.21300552785396576f
#define SS_G5 0.26601171493530273f
__device__ __forceinline__ constexpr float ss_g(int k) {
    return (k == 0 || k == 10) ? SS_G0 : (k == 1 || k == 9) ? SS_G1 : (k == 2 || k == 8) ? SS_G2 : (k == 3 || k == 7) ? SS_G3
         : (k == 4 || k == 6) ? SS_G4 : SS_G5;
}

// 1/x for x in [1e-5, 1e4] (the SSIM denominators: >= C1 or ~C2 > 0): approximate reciprocal + one Newton step, <= 1 ulp, no
// denormal slow path (__frcp_rn compiles to a call with one)
__device__ __forceinline__ float rcp_nr(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return fmaf(r, fmaf(-x, r, 1.0f), r);
}
// keeps a loop-invariant value in its register: without it ptxas re-derives lane / warp / block indices from the special
// registers in every unrolled row (~40 extra instructions per row)
#define SS_KEEP(v) asm volatile("" : "+r"(v))
#define SS_KEEP_PTR(p) asm volatile("" : "+l"(p))      // a finished 64-bit plane pointer, so that p + int is ONE IMAD.WIDE

// cp.async (LDGSTS): 4-byte global -> shared copies that bypass the register file; src_bytes = 0 zero-fills (zero padding,
// rows / columns outside the image).  Each warp keeps SS_NB - 1 input rows in flight: with one row of register prefetch the
// kernel was bound by memory LATENCY (16 warps x 1 row x ~340 B in flight per SM ~ 0.7 TB/s by Little's law).
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int n = valid ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" :: "r"(d), "l"(gmem_src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

// TRAIN: also write the three derivative maps.  MEAN: do not write the map; accumulate it (inside the `crop` border) into
// per-warp partial sums instead.
#ifndef SS_MIN_CTAS
#define SS_MIN_CTAS 1
#endif
#ifndef SS_MIN_CTAS_BWD
#define SS_MIN_CTAS_BWD 2            // backward (two columns per lane): two 8-warp CTAs per SM at <= 128 registers; measured equal
#endif                               // (+-2 %) to 12 warps x 1 CTA at 140 registers and to 8 warps x 1 CTA at 158
template <bool TRAIN, bool MEAN>
__global__ void __launch_bounds__(SS_WARPS * 32, SS_MIN_CTAS)
ssim_fwd_kernel(int H, int W, int rows, float C1, float C2, const float* __restrict__ img1, const float* __restrict__ img2,
                float* __restrict__ ssim_map, float* __restrict__ dm_dmu1, float* __restrict__ dm_dsigma1_sq,
                float* __restrict__ dm_dsigma12, int crop, float* __restrict__ partials)
{
    __shared__ __align__(16) float2 s_row[SS_WARPS][SS_NB][SS_BUFW];  // (u, v) interleaved; a ring of SS_NB rows per warp
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    SS_KEEP(lane); SS_KEEP(warp);
    const size_t plane = (size_t)blockIdx.z * H * W;
    int x0 = (blockIdx.x * SS_WARPS + warp) * 32, y0 = blockIdx.y * rows;
    SS_KEEP(x0); SS_KEEP(y0);
    float local_sum = 0.f;
    if (x0 < W) {
        const int n_out = min(rows, H - y0);
        const int n_in = n_out + 2 * SS_R;
        const int xa = x0 - SS_R + lane, xb = xa + 32;                 // the lane's two input columns (xb only for lane < 10)
        const int px = x0 + lane;
        // column validity etc.: loop-invariant flags in one register
        int flags = ((unsigned)xa < (unsigned)W ? 1 : 0) | ((lane < 2 * SS_R && (unsigned)xb < (unsigned)W) ? 2 : 0) | (px < W ? 4 : 0) |
                    ((px >= crop && px < W - crop) ? 8 : 0) | (lane < 2 * SS_R ? 16 : 0);
        SS_KEEP(flags);
#define ca (flags & 1)
#define cb (flags & 2)
#define cw (flags & 4)
#define cc (flags & 8)
#define lo10 (flags & 16)
        float2* mybuf = &s_row[warp][0][lane];                         // the lane's slot in ring row 0
        f32x2 gg[6];
#pragma unroll
        for (int k = 0; k < 6; k++) gg[k] = pack2(ss_g(k), ss_g(k));
        f32x2 ring_m[11], ring_s[11];                                  // (mu1, mu2), (E[x^2] + E[y^2], E[xy]) after the horizontal pass
        // input element (y0 - 5 + row, xa) of both images.  Addresses are formed from CLAMPED coordinates with 32-bit index math
        // inside the plane (H * W < 2^31, checked by the host): always in bounds, one IMAD + one IMAD.WIDE per copy instead of a
        // 64-bit multiply, two selects and a shift/add pair (the address arithmetic was a quarter of the instructions of a row)
        const float* p1 = img1 + plane;
        const float* p2 = img2 + plane;
        SS_KEEP_PTR(p1); SS_KEEP_PTR(p2);
        int xac = min(max(xa, 0), W - 1), xbc = min(max(xb, 0), W - 1);
        SS_KEEP(xac); SS_KEEP(xbc);
        size_t o = plane + (size_t)y0 * W + px;                        // output element of the current output row
        auto issue_row = [&](int row) {                                // input row `row` of this strip -> ring slot row % SS_NB
            const int y = y0 - SS_R + row;
            const bool yv = (unsigned)y < (unsigned)H && row < n_in;
            const int rowoff = min(max(y, 0), H - 1) * W;
            float2* dst = mybuf + (row & (SS_NB - 1)) * SS_BUFW;
            const bool va_ = yv && ca, vb_ = yv && cb;
            const int ia = rowoff + xac, ib_ = rowoff + xbc;
            cp_async4(&dst->x, p1 + ia, va_);
            cp_async4(&dst->y, p2 + ia, va_);
            if (lo10) {
                cp_async4(&dst[32].x, p1 + ib_, vb_);
                cp_async4(&dst[32].y, p2 + ib_, vb_);
            }
            cp_async_commit();
        };
#pragma unroll
        for (int r = 0; r < SS_NB - 1; r++) issue_row(r);
#pragma unroll 1
        for (int ib = 0; ib < n_in; ib += 11) {
#pragma unroll
            for (int ii = 0; ii < 11; ii++) {
                const int i = ib + ii;
                if (i < n_in) {                                        // warp-uniform
                    cp_async_wait<SS_NB - 2>();                        // this lane's copies of row i have landed ...
                    __syncwarp();                                      // ... and so have the other lanes'; slot (i - 1) is free
                    issue_row(i + SS_NB - 1);                          // (an empty group past the last row keeps the count uniform)
                    const float2* buf = mybuf + (i & (SS_NB - 1)) * SS_BUFW;
                    // horizontal pass: 11 taps, (u,v) pairs straight from shared memory
                    // (two partial sums per moment -- even / odd taps -- halve the dependent-FMA chains: the kernel is bound by
                    // FMA latency x issue, not by memory)
                    // SSIM needs sigma1^2 and sigma2^2 only as their SUM, so x^2 + y^2 is filtered as ONE quantity and travels packed
                    // with xy: two FFMA2 per tap in either pass instead of two FFMA2 + one FFMA, and a 44- instead of 55-register ring.
                    f32x2 hm = 0ull, hs = 0ull, hm1 = 0ull, hs1 = 0ull;        // (mu1, mu2), (E[x^2] + E[y^2], E[xy])
                    // the window is symmetric: taps k and 10 - k share a weight, so they are summed before the weight is applied
#pragma unroll
                    for (int k = 0; k < 5; k++) {
                        const float2 uv = buf[k], wz = buf[10 - k];
                        const f32x2 p = pack2(uv.x, uv.y), q = pack2(wz.x, wz.y);
                        const f32x2 g2 = gg[k];
                        const f32x2 sm = add2(p, q);
                        float t0, t1;
                        unpack2(fma2(p, p, mul2(q, q)), t0, t1);              // (u^2 + w^2, v^2 + z^2)
                        const f32x2 sq = pack2(t0 + t1, fmaf(uv.x, uv.y, wz.x * wz.y));
                        if (k & 1) { hm1 = fma2(g2, sm, hm1); hs1 = fma2(g2, sq, hs1); }
                        else { hm = fma2(g2, sm, hm); hs = fma2(g2, sq, hs); }
                    }
                    {
                        const float2 uv = buf[5];
                        hm1 = fma2(gg[5], pack2(uv.x, uv.y), hm1);
                        hs1 = fma2(gg[5], pack2(fmaf(uv.x, uv.x, uv.y * uv.y), uv.x * uv.y), hs1);
                    }
                    const f32x2 one2 = pack2(1.0f, 1.0f);
                    ring_m[ii] = fma2(hm1, one2, hm); ring_s[ii] = fma2(hs1, one2, hs);
                    if (i >= 2 * SS_R) {
                        // vertical pass over the register ring: the oldest row sits at (ii + 1) % 11
                        f32x2 vm = 0ull, vs = 0ull, vm1 = 0ull, vs1 = 0ull;
#pragma unroll
                        for (int k = 0; k < 11; k++) {
                            const int r = (ii + 1 + k) % 11;
                            const f32x2 g2 = gg[k <= 5 ? k : 10 - k];
                            if (k & 1) { vm1 = fma2(g2, ring_m[r], vm1); vs1 = fma2(g2, ring_s[r], vs1); }
                            else { vm = fma2(g2, ring_m[r], vm); vs = fma2(g2, ring_s[r], vs); }
                        }
                        vm = fma2(vm1, one2, vm); vs = fma2(vs1, one2, vs);
                        if (cw) {
                            float mu1, mu2, ess, e12;
                            unpack2(vm, mu1, mu2); unpack2(vs, ess, e12);
                            const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu1_mu2 = mu1 * mu2;
                            const float musq = mu1_sq + mu2_sq, sigma12 = e12 - mu1_mu2;
                            const float Cc = 2.0f * mu1_mu2 + C1, D = 2.0f * sigma12 + C2;
                            const float A = musq + C1, B = (ess - musq) + C2;   // sigma1^2 + sigma2^2 = E[x^2] + E[y^2] - mu1^2 - mu2^2
                            const float rA = rcp_nr(A), rB = rcp_nr(B), rAB = rA * rB;
                            const float m = Cc * D * rAB;
                            if (MEAN) {
                                const int py = y0 + i - 2 * SS_R;
                                if (cc && py >= crop && py < H - crop) local_sum += m;
                            } else {
                                ssim_map[o] = m;
                            }
                            if (TRAIN) {
                                // d ssim / d mu1, d sigma1^2, d sigma12  (ssim.cu:264-283, the same algebra with 1/A, 1/B factored out)
                                dm_dmu1[o] = 2.0f * rAB * (mu2 * (D - Cc) + mu1 * (Cc * D) * (rB - rA));
                                dm_dsigma1_sq[o] = -m * rB;
                                dm_dsigma12[o] = 2.0f * Cc * rAB;
                            }
                        }
                        o += W;
                    }
                }
            }
        }
        cp_async_wait<0>();
#undef ca
#undef cb
#undef cw
#undef cc
#undef lo10
    }
    if (MEAN) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) local_sum += __shfl_xor_sync(0xFFFFFFFFu, local_sum, o);
        if (lane == 0)
            partials[((size_t)blockIdx.z * gridDim.y + blockIdx.y) * (gridDim.x * SS_WARPS) + blockIdx.x * SS_WARPS + warp] = local_sum;
    }
}

// Fixed-order fp64 sum of the per-warp partial sums -> mean (one CTA; deterministic).
__global__ void __launch_bounds__(256)
ssim_mean_finalize_kernel(const float* __restrict__ partials, long long n, double inv_count, float* __restrict__ out)
{
    __shared__ double s[256];
    double a = 0.0;
    for (long long i = threadIdx.x; i < n; i += 256) a += (double)partials[i];
    s[threadIdx.x] = a;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = (float)(s[0] * inv_count);
}

// Backward: dL/dimg1 = conv(dL dm/dmu1) + 2 img1 conv(dL dm/dsigma1^2) + img2 conv(dL dm/dsigma12)   (ssim.cu:288-366).
// MEAN: dL/dmap is the scalar *grad_scalar * scale inside the crop border and 0 outside: never materialised -- the border is
// zero-filled by the copies and the scalar is applied after the convolutions.
//
// A warp owns a 64-column strip and every lane TWO adjacent output columns: the 12 ring values a lane reads per row and stream
// serve both columns and start at an even index, so they arrive as 6 LDS.128 (+ 6 LDS.64) instead of 2 x 11 scalar loads -- the
// one-column form of this kernel was bound by the shared-memory pipe (ncu: LSU 56 %, `mio_throttle` the top stall).  The copies
// keep the coalesced mapping (lane -> column lane, lane + 32, halo lane + 64).
constexpr int SS_BW2 = 76;           // ring row width in pixels: 64 + 2*SS_R = 74, padded so that every array stays 16-byte aligned
// ring row: float2 (dm/dmu1, dm/dsigma1^2)[SS_BW2] | MEAN: float dm/dsigma12[SS_BW2]; else float2 (dm/dsigma12, dL/dmap)[SS_BW2]
template <bool MEAN> __host__ __device__ constexpr int ss_bwd_row_floats() { return (MEAN ? 3 : 4) * SS_BW2; }
// dynamic shared memory per CTA: per warp a ring of SS_NB input rows + SS_NB rows of (img1, img2) at the 64 output pixels
template <bool MEAN> constexpr size_t ss_bwd_smem() { return (size_t)SS_WARPS_B * SS_NB * (ss_bwd_row_floats<MEAN>() * sizeof(float) + 64 * sizeof(float2)); }

template <bool MEAN>
__global__ void __launch_bounds__(SS_WARPS_B * 32, SS_MIN_CTAS_BWD)
ssim_bwd_kernel(int H, int W, int rows, const float* __restrict__ img1, const float* __restrict__ img2,
                const float* __restrict__ dL_dmap, const float* __restrict__ grad_scalar, float scale, int crop,
                const float* __restrict__ dm_dmu1, const float* __restrict__ dm_dsigma1_sq, const float* __restrict__ dm_dsigma12,
                float* __restrict__ dL_dimg1)
{
    extern __shared__ __align__(16) unsigned char ss_dyn[];
    constexpr int RS = ss_bwd_row_floats<MEAN>();          // floats per ring row
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    SS_KEEP(lane); SS_KEEP(warp);
    float* ring = reinterpret_cast<float*>(ss_dyn) + (size_t)warp * SS_NB * RS;
    float2* pixring = reinterpret_cast<float2*>(ss_dyn + (size_t)SS_WARPS_B * SS_NB * RS * sizeof(float)) + (size_t)warp * SS_NB * 64;
    const size_t plane = (size_t)blockIdx.z * H * W;
    int x0 = (blockIdx.x * SS_WARPS_B + warp) * 64, y0 = blockIdx.y * rows;
    SS_KEEP(x0); SS_KEEP(y0);
    if (x0 >= W) return;
    const float gs = MEAN ? __ldg(grad_scalar) * scale : 1.f;
    const int n_out = min(rows, H - y0);
    const int n_in = n_out + 2 * SS_R;
    // copies: the lane brings in ring columns lane, lane + 32 and (lane < 10) lane + 64 = image columns xa, xa + 32, xa + 64, and
    // the (img1, img2) values of output columns x0 + lane, x0 + 32 + lane
    const int xa = x0 - SS_R + lane, xb = xa + 32, xc = xa + 64;
    const int pa = x0 + lane, pb = pa + 32;
    const int px = x0 + 2 * lane;                           // the lane's two OUTPUT columns: px, px + 1
    // a column contributes if it is inside the image and (MEAN) inside the crop border
    int flags = ((xa >= crop && xa < W - crop) ? 1 : 0) | ((xb >= crop && xb < W - crop) ? 2 : 0) |
                ((lane < 2 * SS_R && xc >= crop && xc < W - crop) ? 4 : 0) | (lane < 2 * SS_R ? 8 : 0) |
                (px < W ? 16 : 0) | (px + 1 < W ? 32 : 0) | (pa < W ? 64 : 0) | (pb < W ? 128 : 0);
    SS_KEEP(flags);
#define ca (flags & 1)
#define cb (flags & 2)
#define cc (flags & 4)
#define lo10 (flags & 8)
#define cw0 (flags & 16)
#define cw1 (flags & 32)
#define cpa (flags & 64)
#define cpb (flags & 128)
    f32x2 gg[6];
#pragma unroll
    for (int k = 0; k < 6; k++) gg[k] = pack2(ss_g(k), ss_g(k));
    const float* q1 = dm_dmu1 + plane;
    const float* q2 = dm_dsigma1_sq + plane;
    const float* q3 = dm_dsigma12 + plane;
    const float* q0 = MEAN ? nullptr : dL_dmap + plane;
    const float* i1 = img1 + plane;
    const float* i2 = img2 + plane;
    SS_KEEP_PTR(q1); SS_KEEP_PTR(q2); SS_KEEP_PTR(q3); SS_KEEP_PTR(i1); SS_KEEP_PTR(i2);
    if (!MEAN) SS_KEEP_PTR(q0);
    // (addresses from clamped coordinates, 32-bit index math inside the plane: see the forward kernel)
    int xac = min(max(xa, 0), W - 1), xbc = min(max(xb, 0), W - 1), xcc = min(max(xc, 0), W - 1), pac = min(pa, W - 1), pbc = min(pb, W - 1);
    SS_KEEP(xac); SS_KEEP(xbc); SS_KEEP(xcc); SS_KEEP(pac); SS_KEEP(pbc);
    // one commit group per row index: the input row `row` of the strip AND the image values the output row finished in the
    // same iteration (row - 10) needs -- so the epilogue never waits on a global load
    auto issue_row = [&](int row) {
        const int y = y0 - SS_R + row;
        const bool yv = y >= crop && y < H - crop && row < n_in;
        const int rowoff = min(max(y, 0), H - 1) * W;
        float* dst = ring + (row & (SS_NB - 1)) * RS;
        const bool va_ = yv && ca, vb_ = yv && cb, vc_ = yv && cc;
        const int ia = rowoff + xac, ib_ = rowoff + xbc, ic_ = rowoff + xcc;
        float* dab = dst + 2 * lane;                                     // float2 (a, b)[lane]
        float* dcd = dst + 2 * SS_BW2 + (MEAN ? lane : 2 * lane);        // float c[lane]  /  float2 (c, dl)[lane]
        constexpr int CS = MEAN ? 1 : 2;                                 // floats per pixel of the second array
        cp_async4(dab, q1 + ia, va_); cp_async4(dab + 1, q2 + ia, va_); cp_async4(dcd, q3 + ia, va_);
        if (!MEAN) cp_async4(dcd + 1, q0 + ia, va_);
        cp_async4(dab + 64, q1 + ib_, vb_); cp_async4(dab + 65, q2 + ib_, vb_); cp_async4(dcd + 32 * CS, q3 + ib_, vb_);
        if (!MEAN) cp_async4(dcd + 32 * CS + 1, q0 + ib_, vb_);
        if (lo10) {
            cp_async4(dab + 128, q1 + ic_, vc_); cp_async4(dab + 129, q2 + ic_, vc_); cp_async4(dcd + 64 * CS, q3 + ic_, vc_);
            if (!MEAN) cp_async4(dcd + 64 * CS + 1, q0 + ic_, vc_);
        }
        const int yo = y0 + row - 2 * SS_R;
        const bool vo = row >= 2 * SS_R && row < n_in;
        const int orow = min(max(yo, 0), H - 1) * W;
        float2* pd = pixring + (row & (SS_NB - 1)) * 64 + lane;
        const bool voa = vo && cpa, vob = vo && cpb;
        cp_async4(&pd->x, i1 + (orow + pac), voa); cp_async4(&pd->y, i2 + (orow + pac), voa);
        cp_async4(&pd[32].x, i1 + (orow + pbc), vob); cp_async4(&pd[32].y, i2 + (orow + pbc), vob);
        cp_async_commit();
    };
    f32x2 ring_ab[2][11];
    float ring_c[2][11];
#pragma unroll
    for (int r = 0; r < SS_NB - 1; r++) issue_row(r);
    size_t o = plane + (size_t)y0 * W + px;
    const f32x2 one2 = pack2(1.0f, 1.0f);
#pragma unroll 1
    for (int ib = 0; ib < n_in; ib += 11) {
#pragma unroll
        for (int ii = 0; ii < 11; ii++) {
            const int i = ib + ii;
            if (i < n_in) {
                cp_async_wait<SS_NB - 2>();
                __syncwarp();
                issue_row(i + SS_NB - 1);
                const float* buf = ring + (i & (SS_NB - 1)) * RS;
                // the lane's 12 ring pixels 2*lane .. 2*lane + 11 (taps 0..10 of column 0, taps 0..10 of column 1 shifted by one)
                f32x2 A[12];
                float Cv[12];
#pragma unroll
                for (int j = 0; j < 6; j++) {
                    const float4 t = reinterpret_cast<const float4*>(buf)[lane + j];
                    float a0 = t.x, b0 = t.y, a1 = t.z, b1 = t.w, c0, c1;
                    if (MEAN) {
                        const float2 u = reinterpret_cast<const float2*>(buf + 2 * SS_BW2)[lane + j];
                        c0 = u.x; c1 = u.y;
                    } else {
                        const float4 u = reinterpret_cast<const float4*>(buf + 2 * SS_BW2)[lane + j];
                        c0 = u.x * u.y; a0 *= u.y; b0 *= u.y; c1 = u.z * u.w; a1 *= u.w; b1 *= u.w;
                    }
                    A[2 * j] = pack2(a0, b0); A[2 * j + 1] = pack2(a1, b1); Cv[2 * j] = c0; Cv[2 * j + 1] = c1;
                }
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    f32x2 hab = 0ull, hab1 = 0ull;
                    float hc = 0.f, hc1 = 0.f;
#pragma unroll
                    for (int k = 0; k < 11; k++) {
                        if (k & 1) { hab1 = fma2(gg[k <= 5 ? k : 10 - k], A[k + c], hab1); hc1 = fmaf(ss_g(k), Cv[k + c], hc1); }
                        else { hab = fma2(gg[k <= 5 ? k : 10 - k], A[k + c], hab); hc = fmaf(ss_g(k), Cv[k + c], hc); }
                    }
                    ring_ab[c][ii] = fma2(hab1, one2, hab); ring_c[c][ii] = hc + hc1;
                }
                if (i >= 2 * SS_R) {
                    const float4 pix = reinterpret_cast<const float4*>(pixring + (i & (SS_NB - 1)) * 64)[lane];   // (img1, img2) of px, px + 1
                    float v[2];
#pragma unroll
                    for (int c = 0; c < 2; c++) {
                        f32x2 vab = 0ull, vab1 = 0ull;
                        float vc = 0.f, vc1 = 0.f;
#pragma unroll
                        for (int k = 0; k < 11; k++) {
                            const int r = (ii + 1 + k) % 11;
                            if (k & 1) { vab1 = fma2(gg[k <= 5 ? k : 10 - k], ring_ab[c][r], vab1); vc1 = fmaf(ss_g(k), ring_c[c][r], vc1); }
                            else { vab = fma2(gg[k <= 5 ? k : 10 - k], ring_ab[c][r], vab); vc = fmaf(ss_g(k), ring_c[c][r], vc); }
                        }
                        vab = fma2(vab1, one2, vab); vc += vc1;
                        float a, b;
                        unpack2(vab, a, b);
                        const float p1 = c ? pix.z : pix.x, p2 = c ? pix.w : pix.y;
                        v[c] = a + p1 * 2.0f * b + p2 * vc;
                    }
                    if (cw0) dL_dimg1[o] = MEAN ? gs * v[0] : v[0];
                    if (cw1) dL_dimg1[o + 1] = MEAN ? gs * v[1] : v[1];
                    o += W;
                }
            }
        }
    }
    cp_async_wait<0>();
#undef ca
#undef cb
#undef cc
#undef lo10
#undef cw0
#undef cw1
#undef cpa
#undef cpb
}

// Rows per warp strip: every strip costs (rows + 10) input rows; the launch runs ceil(strips / resident warps) rounds of
// equal strips, so pick the height that minimises rounds x (rows + 10)  (148 SMs x 2 CTAs x 8 warps resident).
// strip_w: columns per warp strip (32 forward, 64 backward).
static int ss_rows(int B, int CH, int H, int W, int strip_w, int warps, int resident_warps_per_sm) {
    const long long slots = 148LL * resident_warps_per_sm;
    const long long cols = (long long)((W + strip_w * warps - 1) / (strip_w * warps)) * warps * B * CH;   // strips per row band (incl. idle warps)
    long long best_cost = -1; int best = 64;
    for (int rows = 24; rows <= SS_ROWS_MAX; rows++) {
        const long long strips = cols * ((H + rows - 1) / rows);
        const long long cost = ((strips + slots - 1) / slots) * (rows + 2 * SS_R);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = rows; }
    }
    return best < H ? best : (H > 0 ? H : 1);
}

static dim3 ss_grid(int B, int CH, int H, int W, int rows, int strip_w, int warps) {
    return dim3((W + strip_w * warps - 1) / (strip_w * warps), (H + rows - 1) / rows, B * CH);
}
constexpr int SS_FWD_STRIP = 32, SS_BWD_STRIP = 64;
// the mean workspace is sized for the smallest strip height ss_rows() may choose
static dim3 ss_grid_max(int B, int CH, int H, int W) { return ss_grid(B, CH, H, W, H < 24 ? (H > 0 ? H : 1) : 24, SS_FWD_STRIP, SS_WARPS); }

template <bool MEAN>
static cudaError_t ss_launch_bwd(dim3 grid, cudaStream_t st, int H, int W, int rows, const float* img1, const float* img2, const float* dL_dmap,
                                 const float* grad_scalar, float scale, int crop, const float* d1, const float* d2, const float* d3, float* out) {
    static bool attr_set = false;      // per instantiation
    if (!attr_set) {
        const cudaError_t e = cudaFuncSetAttribute(ssim_bwd_kernel<MEAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ss_bwd_smem<MEAN>());
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    ssim_bwd_kernel<MEAN><<<grid, SS_WARPS_B * 32, ss_bwd_smem<MEAN>(), st>>>(H, W, rows, img1, img2, dL_dmap, grad_scalar, scale, crop, d1, d2, d3, out);
    return cudaGetLastError();
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
    if ((long long)B * CH > 65535 || (long long)H * W > 0x7FFFFFFFLL) return SSB_ERR_CAPACITY;   // grid.z; 32-bit index math inside a plane
    const int rows = ss_rows(B, CH, H, W, SS_FWD_STRIP, SS_WARPS, SS_FWD_RESIDENT_WARPS);
    const dim3 grid = ss_grid(B, CH, H, W, rows, SS_FWD_STRIP, SS_WARPS);
    cudaStream_t st = (cudaStream_t)stream_;
    if (dm_dmu1) ssim_fwd_kernel<true, false><<<grid, SS_WARPS * 32, 0, st>>>(H, W, rows, C1, C2, img1, img2, ssim_map, dm_dmu1, dm_dsigma1_sq, dm_dsigma12, 0, nullptr);
    else ssim_fwd_kernel<false, false><<<grid, SS_WARPS * 32, 0, st>>>(H, W, rows, C1, C2, img1, img2, ssim_map, nullptr, nullptr, nullptr, 0, nullptr);
    return ssb_set_cuda_error(cudaGetLastError());
}

int ssb_fused_ssim_backward(int B, int CH, int H, int W, float C1, float C2, const float* img1, const float* img2,
                            const float* dL_dmap, const float* dm_dmu1, const float* dm_dsigma1_sq,
                            const float* dm_dsigma12, float* dL_dimg1, void* stream_) {
    (void)C1; (void)C2;
    if (B < 0 || CH < 0 || H < 0 || W < 0) return SSB_ERR_INVALID;
    if ((size_t)B * CH * H * W == 0) return SSB_OK;
    if (!img1 || !img2 || !dL_dmap || !dm_dmu1 || !dm_dsigma1_sq || !dm_dsigma12 || !dL_dimg1) return SSB_ERR_INVALID;
    if ((long long)B * CH > 65535 || (long long)H * W > 0x7FFFFFFFLL) return SSB_ERR_CAPACITY;   // grid.z; 32-bit index math inside a plane
    const int rows = ss_rows(B, CH, H, W, SS_BWD_STRIP, SS_WARPS_B, SS_BWD_RESIDENT_WARPS);
    return ssb_set_cuda_error(ss_launch_bwd<false>(ss_grid(B, CH, H, W, rows, SS_BWD_STRIP, SS_WARPS_B), (cudaStream_t)stream_, H, W, rows, img1, img2, dL_dmap, nullptr, 0.f, 0,
                                                   dm_dmu1, dm_dsigma1_sq, dm_dsigma12, dL_dimg1));
}

size_t ssb_fused_ssim_mean_workspace_bytes(int B, int CH, int H, int W) {
    const dim3 g = ss_grid_max(B > 0 ? B : 1, CH > 0 ? CH : 1, H > 0 ? H : 1, W > 0 ? W : 1);
    return (size_t)g.x * g.y * g.z * SS_WARPS * sizeof(float);
}

int ssb_fused_ssim_mean_forward(int B, int CH, int H, int W, float C1, float C2, const float* img1, const float* img2, int crop,
                                float* mean_out, float* dm_dmu1, float* dm_dsigma1_sq, float* dm_dsigma12, void* workspace, void* stream_) {
    if (B < 0 || CH < 0 || H < 0 || W < 0 || crop < 0) return SSB_ERR_INVALID;
    if ((size_t)B * CH * H * W == 0 || H <= 2 * crop || W <= 2 * crop) return SSB_ERR_INVALID;
    if (!img1 || !img2 || !mean_out || !workspace) return SSB_ERR_INVALID;
    if ((dm_dmu1 != nullptr) != (dm_dsigma1_sq != nullptr) || (dm_dmu1 != nullptr) != (dm_dsigma12 != nullptr)) return SSB_ERR_INVALID;
    if ((long long)B * CH > 65535 || (long long)H * W > 0x7FFFFFFFLL) return SSB_ERR_CAPACITY;   // grid.z; 32-bit index math inside a plane
    const int rows = ss_rows(B, CH, H, W, SS_FWD_STRIP, SS_WARPS, SS_FWD_RESIDENT_WARPS);
    const dim3 grid = ss_grid(B, CH, H, W, rows, SS_FWD_STRIP, SS_WARPS);
    cudaStream_t st = (cudaStream_t)stream_;
    float* partials = (float*)workspace;
    if (dm_dmu1) ssim_fwd_kernel<true, true><<<grid, SS_WARPS * 32, 0, st>>>(H, W, rows, C1, C2, img1, img2, nullptr, dm_dmu1, dm_dsigma1_sq, dm_dsigma12, crop, partials);
    else ssim_fwd_kernel<false, true><<<grid, SS_WARPS * 32, 0, st>>>(H, W, rows, C1, C2, img1, img2, nullptr, nullptr, nullptr, nullptr, crop, partials);
    const double count = (double)B * CH * (double)(H - 2 * crop) * (double)(W - 2 * crop);
    ssim_mean_finalize_kernel<<<1, 256, 0, st>>>(partials, (long long)grid.x * grid.y * grid.z * SS_WARPS, 1.0 / count, mean_out);
    return ssb_set_cuda_error(cudaGetLastError());
}

int ssb_fused_ssim_mean_backward(int B, int CH, int H, int W, const float* img1, const float* img2, const float* grad_mean, int crop,
                                 const float* dm_dmu1, const float* dm_dsigma1_sq, const float* dm_dsigma12, float* dL_dimg1, void* stream_) {
    if (B < 0 || CH < 0 || H < 0 || W < 0 || crop < 0) return SSB_ERR_INVALID;
    if ((size_t)B * CH * H * W == 0 || H <= 2 * crop || W <= 2 * crop) return SSB_ERR_INVALID;
    if (!img1 || !img2 || !grad_mean || !dm_dmu1 || !dm_dsigma1_sq || !dm_dsigma12 || !dL_dimg1) return SSB_ERR_INVALID;
    if ((long long)B * CH > 65535 || (long long)H * W > 0x7FFFFFFFLL) return SSB_ERR_CAPACITY;   // grid.z; 32-bit index math inside a plane
    const double count = (double)B * CH * (double)(H - 2 * crop) * (double)(W - 2 * crop);
    const int rows = ss_rows(B, CH, H, W, SS_BWD_STRIP, SS_WARPS_B, SS_BWD_RESIDENT_WARPS);
    return ssb_set_cuda_error(ss_launch_bwd<true>(ss_grid(B, CH, H, W, rows, SS_BWD_STRIP, SS_WARPS_B), (cudaStream_t)stream_, H, W, rows, img1, img2, nullptr, grad_mean,
                                                  (float)(1.0 / count), crop, dm_dmu1, dm_dsigma1_sq, dm_dsigma12, dL_dimg1));
}

}  // extern "C"
