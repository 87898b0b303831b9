// skelsplat_b200 -- shared device code for the sm_100a rasteriser / optimiser kernels.
//
// Per-Gaussian math of the SkelSplat rasteriser (EWA projection, tile rectangle,
// (tile|depth) keys, alpha evaluation, and the backward chain), written once and
// used by both the dense drop-in rasteriser (raster_dense.cu) and the fused
// per-frame optimiser (optimizer.cu).
//
// Bit-exactness contract: tile keys / sort order / ranges must equal the reference
// CUDA rasteriser's on the same inputs.  nvcc's FMA contraction of the reference's
// expressions is context dependent, so every floating-point operation on the path
// that decides a tile rectangle or a depth key is written with explicit rounding
// intrinsics (__fmaf_rn / __fmul_rn / __fadd_rn ...), in the operation order the
// reference compiles to for sm_100a (reference semantics: RAST/cuda_rasterizer/
// forward.cu:74-273, auxiliary.h:40-89, rasterizer_impl.cu:70-138; RAST =
// submodules/diff-gaussian-rasterization-h36m of the reference).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ssb {

constexpr int TILE = 16;            // BLOCK_X == BLOCK_Y (RAST/cuda_rasterizer/config.h:16-17)
constexpr float NEAR_Z = 0.2f;      // in_frustum, auxiliary.h:166
constexpr float DILATION = 0.3f;    // h_var, forward.cu:219
constexpr float ALPHA_MAX = 0.99f;  // forward.cu:362
constexpr float ALPHA_MIN = 1.0f / 255.0f;
constexpr float T_EPS = 0.0001f;    // forward.cu:366

// ---- exact-order helpers ------------------------------------------------------------------
// a*x + b*y + c*z + d  ->  t=b*y; t=fma(a,x,t); t=fma(c,z,t); d+t
__device__ __forceinline__ float affine3(float a, float x, float b, float y, float c, float z, float d) {
    float t = __fmul_rn(b, y);
    t = __fmaf_rn(a, x, t);
    t = __fmaf_rn(c, z, t);
    return __fadd_rn(d, t);
}
// fma(a2,b2, fma(a0,b0, a1*b1))
__device__ __forceinline__ float dot3(float a0, float b0, float a1, float b1, float a2, float b2) {
    float t = __fmul_rn(a1, b1);
    t = __fmaf_rn(a0, b0, t);
    return __fmaf_rn(a2, b2, t);
}
// ((v + 1.0) * S - 1.0) * 0.5 in fp64 with one fma, rounded to fp32 (auxiliary.h:40-43)
__device__ __forceinline__ float ndc2pix(float v, int S) {
    double d = __fma_rn(__dadd_rn((double)v, 1.0), (double)S, -1.0);
    return __double2float_rn(__dmul_rn(d, 0.5));
}

struct Rect { uint32_t x0, y0, x1, y1; };

// getRect, auxiliary.h:45-55
__device__ __forceinline__ Rect tile_rect(float px, float py, int radius, int gx, int gy) {
    const float rf = (float)radius;
    Rect r;
    r.x0 = (uint32_t)min(gx, max(0, (int)__fmul_rn(__fsub_rn(px, rf), 0.0625f)));
    r.y0 = (uint32_t)min(gy, max(0, (int)__fmul_rn(__fsub_rn(py, rf), 0.0625f)));
    r.x1 = (uint32_t)min(gx, max(0, (int)__fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(px, rf), 16.0f), -1.0f), 0.0625f)));
    r.y1 = (uint32_t)min(gy, max(0, (int)__fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(py, rf), 16.0f), -1.0f), 0.0625f)));
    return r;
}

// Sigma = (S R)^T (S R) with the quaternion used UN-normalised (forward.cu:114-150)
__device__ __forceinline__ void cov3d_from_scale_rot(float sx0, float sy0, float sz0, float mod,
                                                      float r, float x, float y, float z, float* cov) {
    const float sx = __fmul_rn(mod, sx0), sy = __fmul_rn(mod, sy0), sz = __fmul_rn(mod, sz0);
    const float yy = __fmul_rn(y, y), zz = __fmul_rn(z, z);
    const float rz = __fmul_rn(r, z), xz = __fmul_rn(x, z), rx = __fmul_rn(r, x);
    // operation order of the reference as compiled for sm_100a (ptxas fuses one product of each a*b +- c*d)
    float h;
    h = __fadd_rn(yy, zz);               const float R00 = __fsub_rn(1.0f, __fadd_rn(h, h));
    h = __fmaf_rn(x, y, -rz);            const float R01 = __fadd_rn(h, h);       // xy - rz
    h = __fmaf_rn(r, y, xz);             const float R02 = __fadd_rn(h, h);       // xz + ry
    h = __fmaf_rn(x, y, rz);             const float R10 = __fadd_rn(h, h);       // xy + rz
    h = __fmaf_rn(x, x, zz);             const float R11 = __fsub_rn(1.0f, __fadd_rn(h, h));
    h = __fmaf_rn(y, z, -rx);            const float R12 = __fadd_rn(h, h);       // yz - rx
    h = __fmaf_rn(-r, y, xz);            const float R20 = __fadd_rn(h, h);       // xz - ry
    h = __fmaf_rn(y, z, rx);             const float R21 = __fadd_rn(h, h);       // yz + rx
    h = __fmaf_rn(x, x, yy);             const float R22 = __fsub_rn(1.0f, __fadd_rn(h, h));
    const float A0 = __fmul_rn(sx, R00), A1 = __fmul_rn(sy, R01), A2 = __fmul_rn(sz, R02);
    const float B0 = __fmul_rn(sx, R10), B1 = __fmul_rn(sy, R11), B2 = __fmul_rn(sz, R12);
    const float C0 = __fmul_rn(sx, R20), C1 = __fmul_rn(sy, R21), C2 = __fmul_rn(sz, R22);
    cov[0] = dot3(A0, A0, A1, A1, A2, A2);
    cov[1] = dot3(B0, A0, B1, A1, B2, A2);
    cov[2] = dot3(C0, A0, C1, A1, C2, A2);
    cov[3] = dot3(B0, B0, B1, B1, B2, B2);
    cov[4] = dot3(C0, B0, C1, B1, C2, B2);
    cov[5] = dot3(C0, C0, C1, C1, C2, C2);
}

// Everything preprocessCUDA (forward.cu:153-273) leaves behind for one Gaussian.
struct Splat {
    float depth;         // view-space z (the key's low 32 bits)
    float px, py;        // pixel-space mean
    float conx, cony, conz, opac;   // conic + opacity (opacity * h_convolution_scaling)
    int radius;          // 0 => culled
    Rect rect;
    uint32_t tiles;      // tiles_touched
};

// view / proj: 16 floats, element (r,c) of the true matrix at m[4*c+r] (auxiliary.h:70-89).
__device__ __forceinline__ Splat project_gaussian(float mx, float my, float mz, const float* cov3D, float opacity,
                                                   const float* __restrict__ view, const float* __restrict__ proj,
                                                   int W, int H, float tan_fovx, float tan_fovy,
                                                   float focal_x, float focal_y, bool antialiasing) {
    Splat s;
    s.radius = 0; s.tiles = 0; s.rect = Rect{0, 0, 0, 0};
    s.depth = 0.f; s.px = s.py = 0.f; s.conx = s.cony = s.conz = s.opac = 0.f;
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    const float depth = affine3(view[2], mx, view[6], my, view[10], mz, view[14]);
    if (depth <= NEAR_Z) return s;
    const float hx = affine3(proj[0], mx, proj[4], my, proj[8], mz, proj[12]);
    const float hy = affine3(proj[1], mx, proj[5], my, proj[9], mz, proj[13]);
    const float hw = affine3(proj[3], mx, proj[7], my, proj[11], mz, proj[15]);
    const float p_w = __frcp_rn(__fadd_rn(hw, 0.0000001f));
    const float projx = __fmul_rn(hx, p_w), projy = __fmul_rn(hy, p_w);
    // computeCov2D (forward.cu:74-109)
    const float tx = affine3(view[0], mx, view[4], my, view[8], mz, view[12]);
    const float ty = affine3(view[1], mx, view[5], my, view[9], mz, view[13]);
    const float tz = depth;
    const float limx = __fmul_rn(tan_fovx, 1.3f), limy = __fmul_rn(tan_fovy, 1.3f);
    const float txtz = __fdiv_rn(tx, tz), tytz = __fdiv_rn(ty, tz);
    const float cx = fminf(limx, fmaxf(-limx, txtz));
    const float cy = fminf(limy, fmaxf(-limy, tytz));
    const float tz2 = __fmul_rn(tz, tz);
    const float J00 = __fdiv_rn(focal_x, tz);
    const float J02 = __fdiv_rn(__fmul_rn(focal_x, __fmul_rn(cx, -tz)), tz2);
    const float J11 = __fdiv_rn(focal_y, tz);
    const float J12 = __fdiv_rn(__fmul_rn(focal_y, __fmul_rn(cy, -tz)), tz2);
    const float a0 = __fmaf_rn(view[2], J02, __fmul_rn(view[0], J00));
    const float a1 = __fmaf_rn(view[6], J02, __fmul_rn(view[4], J00));
    const float a2 = __fmaf_rn(J02, view[10], __fmul_rn(view[8], J00));
    const float b0 = __fmaf_rn(view[2], J12, __fmul_rn(J11, view[1]));
    const float b1 = __fmaf_rn(view[6], J12, __fmul_rn(J11, view[5]));
    const float b2 = __fmaf_rn(J12, view[10], __fmul_rn(J11, view[9]));
    const float c0 = cov3D[0], c1 = cov3D[1], c2 = cov3D[2], c3 = cov3D[3], c4 = cov3D[4], c5 = cov3D[5];
    const float va0 = dot3(a0, c0, a1, c1, a2, c2), vb0 = dot3(b0, c0, b1, c1, b2, c2);
    const float va1 = dot3(a0, c1, a1, c3, a2, c4), vb1 = dot3(b0, c1, b1, c3, b2, c4);
    const float va2 = dot3(a0, c2, a1, c4, a2, c5), vb2 = dot3(b0, c2, b1, c4, b2, c5);
    float cov_x = dot3(a0, va0, a1, va1, a2, va2);
    const float cov_y = dot3(a0, vb0, a1, vb1, a2, vb2);
    float cov_z = dot3(b0, vb0, b1, vb1, b2, vb2);
    const float cyy = __fmul_rn(cov_y, cov_y);
    const float det_cov = __fmaf_rn(cov_x, cov_z, -cyy);
    cov_x = __fadd_rn(cov_x, DILATION);
    cov_z = __fadd_rn(cov_z, DILATION);
    const float det = __fmaf_rn(cov_x, cov_z, -cyy);
    float h_scaling = 1.0f;
    if (antialiasing) h_scaling = __fsqrt_rn(fmaxf(0.000025f, __fdiv_rn(det_cov, det)));
    if (det == 0.0f) return s;
    const float det_inv = __frcp_rn(det);
    const float mid = __fmul_rn(__fadd_rn(cov_x, cov_z), 0.5f);
    const float root = __fsqrt_rn(fmaxf(__fmaf_rn(mid, mid, -det), 0.1f));
    const float lambda1 = __fadd_rn(mid, root), lambda2 = __fsub_rn(mid, root);
    const float my_radius = ceilf(__fmul_rn(__fsqrt_rn(fmaxf(lambda1, lambda2)), 3.0f));
    const float pix_x = ndc2pix(projx, W), pix_y = ndc2pix(projy, H);
    const int rad = (int)my_radius;
    const Rect r = tile_rect(pix_x, pix_y, rad, gx, gy);
    const uint32_t tiles = (r.x1 - r.x0) * (r.y1 - r.y0);
    if (tiles == 0) return s;
    s.depth = depth; s.px = pix_x; s.py = pix_y;
    s.conx = __fmul_rn(cov_z, det_inv); s.cony = __fmul_rn(det_inv, -cov_y); s.conz = __fmul_rn(cov_x, det_inv);
    s.opac = __fmul_rn(h_scaling, opacity);
    s.radius = rad; s.rect = r; s.tiles = tiles;
    return s;
}

// One (pixel, Gaussian) pair, forward.cu:352-364.  Returns false when the pair is skipped.
__device__ __forceinline__ bool pair_alpha(float gpx, float gpy, float conx, float cony, float conz, float opac,
                                           float pxf, float pyf, float& dx, float& dy, float& G, float& alpha) {
    dx = __fsub_rn(gpx, pxf);
    dy = __fsub_rn(gpy, pyf);
    float t = __fmul_rn(dy, __fmul_rn(dy, conz));
    t = __fmaf_rn(dx, __fmul_rn(dx, conx), t);
    const float power = __fmaf_rn(t, -0.5f, -__fmul_rn(dy, __fmul_rn(dx, cony)));
    if (power > 0.0f) return false;
    // exact early-out: opac <= 1 and power < -5.55 imply opac*exp(power) <= 0.003888 < 1/255, the pair is skipped anyway
    if (power < -5.55f && opac <= 1.0f) return false;
    G = expf(power);
    alpha = fminf(ALPHA_MAX, __fmul_rn(opac, G));
    return !(alpha < ALPHA_MIN);
}

// ---- in-shared-memory (tile|depth) sort ------------------------------------------------------
// Bitonic sort of n (power of two) 64-bit keys with 32-bit payloads by ALL threads of the CTA.
// Ties are broken on the payload's low bits (the emission index), which reproduces the order of
// the reference's stable LSD radix sort (rasterizer_impl.cu:303-311).  Padding entries must carry
// key = ~0ull.
__device__ __forceinline__ void bitonic_sort_cta(uint64_t* keys, uint32_t* vals, int n) {
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const uint64_t ka = keys[i], kb = keys[ixj];
                    const uint32_t va = vals[i], vb = vals[ixj];
                    const bool a_gt_b = (ka > kb) || (ka == kb && va > vb);
                    const bool up = ((i & k) == 0);
                    if (a_gt_b == up) {
                        keys[i] = kb; keys[ixj] = ka;
                        vals[i] = vb; vals[ixj] = va;
                    }
                }
            }
            __syncthreads();
        }
    }
}

// getHigherMsb (rasterizer_impl.cu:35-50); only used to validate that tile ids fit the sorted bits.
__host__ __device__ inline uint32_t higher_msb(uint32_t n) {
    uint32_t msb = sizeof(n) * 4, step = msb;
    while (step > 1) {
        step /= 2;
        if (n >> msb) msb += step; else msb -= step;
    }
    if (n >> msb) msb++;
    return msb;
}

// ---- backward of the per-Gaussian chain --------------------------------------------------------
// computeCov2DCUDA (backward.cu:147-326, antialiasing == false) + preprocessCUDA backward
// (backward.cu:398-449) + computeCov3D backward (backward.cu:330-393), for one Gaussian.
// Inputs: gradient w.r.t. pixel-space mean (already scaled by 0.5*W, 0.5*H as renderCUDA does),
// conic (x, y, w entries) and inverse depth.  has_scale_rot == false: cov3D was precomputed.
struct SplatGrad {
    float dmean[3];
    float dcov[6];
    float dscale[3];
    float drot[4];
};

__device__ __forceinline__ SplatGrad gaussian_backward(
    float mx, float my, float mz, const float* cov3D, bool has_scale_rot,
    float s0, float s1, float s2, float mod, float qr, float qx, float qy, float qz,
    const float* __restrict__ view, const float* __restrict__ proj,
    float h_x, float h_y, float tan_fovx, float tan_fovy,
    float d2x, float d2y, float dconx, float dcony, float dconz, float dinvdepth, bool has_invdepth)
{
    SplatGrad g;
    float tx = view[0] * mx + view[4] * my + view[8] * mz + view[12];
    float ty = view[1] * mx + view[5] * my + view[9] * mz + view[13];
    const float tz = view[2] * mx + view[6] * my + view[10] * mz + view[14];
    const float limx = 1.3f * tan_fovx, limy = 1.3f * tan_fovy;
    const float txtz = tx / tz, tytz = ty / tz;
    tx = fminf(limx, fmaxf(-limx, txtz)) * tz;
    ty = fminf(limy, fmaxf(-limy, tytz)) * tz;
    const float x_grad_mul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
    const float y_grad_mul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
    // T = W*J has two non-zero columns a (x) and b (y); W rows are view[0,4,8], view[1,5,9], view[2,6,10]
    const float J00 = h_x / tz, J02 = -(h_x * tx) / (tz * tz);
    const float J11 = h_y / tz, J12 = -(h_y * ty) / (tz * tz);
    const float a0 = view[0] * J00 + view[2] * J02, a1 = view[4] * J00 + view[6] * J02, a2 = view[8] * J00 + view[10] * J02;
    const float b0 = view[1] * J11 + view[2] * J12, b1 = view[5] * J11 + view[6] * J12, b2 = view[9] * J11 + view[10] * J12;
    const float c0 = cov3D[0], c1 = cov3D[1], c2 = cov3D[2], c3 = cov3D[3], c4 = cov3D[4], c5 = cov3D[5];
    // Vrk*a, Vrk*b
    const float va0 = c0 * a0 + c1 * a1 + c2 * a2, va1 = c1 * a0 + c3 * a1 + c4 * a2, va2 = c2 * a0 + c4 * a1 + c5 * a2;
    const float vb0 = c0 * b0 + c1 * b1 + c2 * b2, vb1 = c1 * b0 + c3 * b1 + c4 * b2, vb2 = c2 * b0 + c4 * b1 + c5 * b2;
    const float c_xx = a0 * va0 + a1 * va1 + a2 * va2 + DILATION;
    const float c_xy = a0 * vb0 + a1 * vb1 + a2 * vb2;
    const float c_yy = b0 * vb0 + b1 * vb1 + b2 * vb2 + DILATION;
    const float denom = c_xx * c_yy - c_xy * c_xy;
    const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
    float dxx = 0.f, dxy = 0.f, dyy = 0.f;
    if (denom2inv != 0) {
        dxx = denom2inv * (-c_yy * c_yy * dconx + 2 * c_xy * c_yy * dcony + (denom - c_xx * c_yy) * dconz);
        dyy = denom2inv * (-c_xx * c_xx * dconz + 2 * c_xx * c_xy * dcony + (denom - c_xx * c_yy) * dconx);
        dxy = denom2inv * 2 * (c_xy * c_yy * dconx - (denom + 2 * c_xy * c_xy) * dcony + c_xx * c_xy * dconz);
        g.dcov[0] = a0 * a0 * dxx + a0 * b0 * dxy + b0 * b0 * dyy;
        g.dcov[3] = a1 * a1 * dxx + a1 * b1 * dxy + b1 * b1 * dyy;
        g.dcov[5] = a2 * a2 * dxx + a2 * b2 * dxy + b2 * b2 * dyy;
        g.dcov[1] = 2 * a0 * a1 * dxx + (a0 * b1 + a1 * b0) * dxy + 2 * b0 * b1 * dyy;
        g.dcov[2] = 2 * a0 * a2 * dxx + (a0 * b2 + a2 * b0) * dxy + 2 * b0 * b2 * dyy;
        g.dcov[4] = 2 * a2 * a1 * dxx + (a1 * b2 + a2 * b1) * dxy + 2 * b1 * b2 * dyy;
    } else {
#pragma unroll
        for (int i = 0; i < 6; i++) g.dcov[i] = 0.f;
    }
    // dL/dT (upper 2x3), then dL/dJ, then dL/dt
    const float dT00 = 2 * va0 * dxx + vb0 * dxy, dT01 = 2 * va1 * dxx + vb1 * dxy, dT02 = 2 * va2 * dxx + vb2 * dxy;
    const float dT10 = 2 * vb0 * dyy + va0 * dxy, dT11 = 2 * vb1 * dyy + va1 * dxy, dT12 = 2 * vb2 * dyy + va2 * dxy;
    const float dJ00 = view[0] * dT00 + view[4] * dT01 + view[8] * dT02;
    const float dJ02 = view[2] * dT00 + view[6] * dT01 + view[10] * dT02;
    const float dJ11 = view[1] * dT10 + view[5] * dT11 + view[9] * dT12;
    const float dJ12 = view[2] * dT10 + view[6] * dT11 + view[10] * dT12;
    const float itz = 1.f / tz, itz2 = itz * itz, itz3 = itz2 * itz;
    const float dtx = x_grad_mul * -h_x * itz2 * dJ02;
    const float dty = y_grad_mul * -h_y * itz2 * dJ12;
    float dtz = -h_x * itz2 * dJ00 - h_y * itz2 * dJ11 + (2 * h_x * tx) * itz3 * dJ02 + (2 * h_y * ty) * itz3 * dJ12;
    if (has_invdepth) dtz -= dinvdepth / (tz * tz);
    float gmx = view[0] * dtx + view[1] * dty + view[2] * dtz;
    float gmy = view[4] * dtx + view[5] * dty + view[6] * dtz;
    float gmz = view[8] * dtx + view[9] * dty + view[10] * dtz;
    // projection part (backward.cu:424-440)
    const float m_w = 1.0f / ((proj[3] * mx + proj[7] * my + proj[11] * mz + proj[15]) + 0.0000001f);
    const float mul1 = (proj[0] * mx + proj[4] * my + proj[8] * mz + proj[12]) * m_w * m_w;
    const float mul2 = (proj[1] * mx + proj[5] * my + proj[9] * mz + proj[13]) * m_w * m_w;
    gmx += (proj[0] * m_w - proj[3] * mul1) * d2x + (proj[1] * m_w - proj[3] * mul2) * d2y;
    gmy += (proj[4] * m_w - proj[7] * mul1) * d2x + (proj[5] * m_w - proj[7] * mul2) * d2y;
    gmz += (proj[8] * m_w - proj[11] * mul1) * d2x + (proj[9] * m_w - proj[11] * mul2) * d2y;
    g.dmean[0] = gmx; g.dmean[1] = gmy; g.dmean[2] = gmz;
#pragma unroll
    for (int i = 0; i < 3; i++) g.dscale[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 4; i++) g.drot[i] = 0.f;
    if (has_scale_rot) {
        // rows of the rotation matrix in the reference's storage: Rc[c][r]
        const float r = qr, x = qx, y = qy, z = qz;
        const float R[3][3] = {
            {1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y)},
            {2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x)},
            {2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)}};
        const float s[3] = {mod * s0, mod * s1, mod * s2};
        const float dS[3][3] = {{g.dcov[0], 0.5f * g.dcov[1], 0.5f * g.dcov[2]},
                                {0.5f * g.dcov[1], g.dcov[3], 0.5f * g.dcov[4]},
                                {0.5f * g.dcov[2], 0.5f * g.dcov[4], g.dcov[5]}};
        float dMt[3][3];   // dMt[c][r] = dM[r][c], dM[c][r] = 2 * sum_k (s[r] R[k][r]) dS[c][k]
#pragma unroll
        for (int c = 0; c < 3; c++)
#pragma unroll
            for (int rr = 0; rr < 3; rr++) {
                float a = 0.f;
#pragma unroll
                for (int k = 0; k < 3; k++) a += (s[c] * R[k][c]) * dS[rr][k];
                dMt[c][rr] = 2.0f * a;
            }
#pragma unroll
        for (int c = 0; c < 3; c++) {
            float a = 0.f;
#pragma unroll
            for (int k = 0; k < 3; k++) a += R[k][c] * dMt[c][k];
            g.dscale[c] = a;
        }
#pragma unroll
        for (int c = 0; c < 3; c++)
#pragma unroll
            for (int k = 0; k < 3; k++) dMt[c][k] *= s[c];
        g.drot[0] = 2 * z * (dMt[0][1] - dMt[1][0]) + 2 * y * (dMt[2][0] - dMt[0][2]) + 2 * x * (dMt[1][2] - dMt[2][1]);
        g.drot[1] = 2 * y * (dMt[1][0] + dMt[0][1]) + 2 * z * (dMt[2][0] + dMt[0][2]) + 2 * r * (dMt[1][2] - dMt[2][1]) - 4 * x * (dMt[2][2] + dMt[1][1]);
        g.drot[2] = 2 * x * (dMt[1][0] + dMt[0][1]) + 2 * r * (dMt[2][0] - dMt[0][2]) + 2 * z * (dMt[1][2] + dMt[2][1]) - 4 * y * (dMt[2][2] + dMt[0][0]);
        g.drot[3] = 2 * r * (dMt[0][1] - dMt[1][0]) + 2 * x * (dMt[2][0] + dMt[0][2]) + 2 * y * (dMt[1][2] + dMt[2][1]) - 4 * z * (dMt[1][1] + dMt[0][0]);
    }
    return g;
}

}  // namespace ssb
