// Per-frame setup on the GPU (SURVEY.md section 8 "next" rows f-3 and f-1), sm_100a.
//
//  * dlt_kernel           batched DLT triangulation: the producer of the initial guess
//                         (reference: triangulation.py:122-150, one numpy SVD per joint on the CPU).
//  * roi_rect_kernel /    pseudo-ground-truth heatmaps as ROI patches, batched over frames x views x joints
//    roi_fill_kernel      (reference: utils/general_utils.py:175-304, V*J cupy gaussian_filter calls on megapixel
//                         images with .item() syncs per frame).
// Both are embarrassingly parallel and tiny per item; they exist so that initial guess -> heatmaps -> optimisation
// runs as one GPU pipeline with no per-frame host work.
#include "api_internal.h"

namespace ssb {

// ------------------------------------------------------------------------------------------ DLT
constexpr int DLT_MAXV = 16;

// Smallest-eigenvalue eigenvector of a symmetric 4x4 matrix by cyclic Jacobi rotations (fp64).
__device__ void smallest_eigvec4(double M[4][4], double out[4]) {
    double Vm[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
    for (int sweep = 0; sweep < 30; sweep++) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < 4; i++) {
            diag += M[i][i] * M[i][i];
            for (int j = i + 1; j < 4; j++) off += M[i][j] * M[i][j];
        }
        if (off <= 1e-60 * diag || off == 0.0) break;
        for (int pi = 0; pi < 3; pi++)
            for (int qi = pi + 1; qi < 4; qi++) {
                const double apq = M[pi][qi];
                if (apq == 0.0) continue;
                const double theta = (M[qi][qi] - M[pi][pi]) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
                for (int k = 0; k < 4; k++) {            // M <- M J
                    const double mkp = M[k][pi], mkq = M[k][qi];
                    M[k][pi] = c * mkp - sn * mkq;
                    M[k][qi] = sn * mkp + c * mkq;
                }
                for (int k = 0; k < 4; k++) {            // M <- J^T M
                    const double mpk = M[pi][k], mqk = M[qi][k];
                    M[pi][k] = c * mpk - sn * mqk;
                    M[qi][k] = sn * mpk + c * mqk;
                }
                for (int k = 0; k < 4; k++) {
                    const double vkp = Vm[k][pi], vkq = Vm[k][qi];
                    Vm[k][pi] = c * vkp - sn * vkq;
                    Vm[k][qi] = sn * vkp + c * vkq;
                }
            }
    }
    int best = 0;
    for (int i = 1; i < 4; i++) if (M[i][i] < M[best][best]) best = i;
    for (int k = 0; k < 4; k++) out[k] = Vm[k][best];
}

// One thread per (frame, joint).  P: [V,3,4] fp64 projection matrices K[R|t]; poses_2d: [F,V,J,2] fp64.
__global__ void dlt_kernel(int F, int V, int J, const double* __restrict__ P, const double* __restrict__ poses_2d,
                           double* __restrict__ out /* [F,J,3] */)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= F * J) return;
    const int f = i / J, j = i % J;
    double M[4][4];
    for (int a = 0; a < 4; a++) for (int b = 0; b < 4; b++) M[a][b] = 0.0;
    for (int v = 0; v < V; v++) {
        const double* Pv = P + (size_t)v * 12;
        const double x = poses_2d[(((size_t)f * V + v) * J + j) * 2], y = poses_2d[(((size_t)f * V + v) * J + j) * 2 + 1];
        double r0[4], r1[4];
        for (int k = 0; k < 4; k++) { r0[k] = x * Pv[8 + k] - Pv[k]; r1[k] = y * Pv[8 + k] - Pv[4 + k]; }   // triangulation.py:128-129
        for (int a = 0; a < 4; a++) for (int b = a; b < 4; b++) M[a][b] += r0[a] * r0[b] + r1[a] * r1[b];
    }
    for (int a = 0; a < 4; a++) for (int b = 0; b < a; b++) M[a][b] = M[b][a];
    double x4[4];
    smallest_eigvec4(M, x4);
    out[(size_t)i * 3] = x4[0] / x4[3]; out[(size_t)i * 3 + 1] = x4[1] / x4[3]; out[(size_t)i * 3 + 2] = x4[2] / x4[3];
}

// ------------------------------------------------------------------------------------------ heatmap ROIs
// sigma of the (view, joint) heatmap from the INITIAL Gaussian, restating utils/general_utils.py:199-265 (fp32 torch ops on
// the reference's GPU) in fp64 with a fixed operation order, rounded to fp32 at the end -- the form in which the host
// specification and this kernel agree exactly on every integer window:
// T = R_w2c @ J_rows (not the rasteriser's EWA form), cov = T^T Sigma^T T, +0.3 on the diagonal,
// lambda = mid +- sqrt(max(0.1, mid^2 - det)); sigma_y = sqrt(lambda1) (axis 0), sigma_x = sqrt(lambda2) (axis 1).
struct RoiParams {
    int F, V, J;
    const float* xyz; const float* scaling_raw; const float* rotation_raw;   // [F,J,3] [F,J,3] [F,J,4]
    const float* poses_2d;                                                    // [F,V,J,2]
    ssb_cameras cams;
    float scaling_modifier;
};

__device__ __forceinline__ int gauss_radius(float sigma) { return (int)(4.0 * (double)sigma + 0.5); }   // scipy: int(truncate * sd + 0.5), python floats

// fp64 without FMA contraction: the host specification (skelsplat_b200/heatmaps.py:heatmap_sigmas) performs the same IEEE
// operations in the same order in numpy float64, so both sides produce the same sigma (and hence the same integer window).
__device__ __forceinline__ double dm(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double da(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double ds(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dd(double a, double b) { return __ddiv_rn(a, b); }

__global__ void roi_rect_kernel(RoiParams p, int* __restrict__ roi_rect, float* __restrict__ roi_sigma, int* __restrict__ roi_center,
                                long long* __restrict__ roi_size)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.F * p.V * p.J) return;
    const int j = i % p.J, v = (i / p.J) % p.V, f = i / (p.J * p.V);
    const int W = p.cams.dims ? p.cams.dims[2 * v] : p.cams.W0, H = p.cams.dims ? p.cams.dims[2 * v + 1] : p.cams.H0;
    const double tfx = (double)(p.cams.tanfov ? p.cams.tanfov[2 * v] : p.cams.tanfovx0), tfy = (double)(p.cams.tanfov ? p.cams.tanfov[2 * v + 1] : p.cams.tanfovy0);
    double vm[16];                                           // stored transposed: W2C(r,c) = vm[4*c + r]
    for (int k = 0; k < 16; k++) vm[k] = (double)p.cams.viewmatrix[16 * v + k];
    const size_t gj = (size_t)f * p.J + j;
    const double mx = (double)p.xyz[3 * gj], my = (double)p.xyz[3 * gj + 1], mz = (double)p.xyz[3 * gj + 2];
    // Sigma = (R S)(R S)^T with the NORMALISED quaternion (build_rotation normalises, general_utils.py:87-108)
    const double q0 = (double)p.rotation_raw[4 * gj], q1 = (double)p.rotation_raw[4 * gj + 1], q2 = (double)p.rotation_raw[4 * gj + 2], q3 = (double)p.rotation_raw[4 * gj + 3];
    const double qn = __dsqrt_rn(da(da(da(dm(q0, q0), dm(q1, q1)), dm(q2, q2)), dm(q3, q3)));
    const double r = dd(q0, qn), x = dd(q1, qn), y = dd(q2, qn), z = dd(q3, qn);
    const double R[3][3] = {{ds(1.0, dm(2.0, da(dm(y, y), dm(z, z)))), dm(2.0, ds(dm(x, y), dm(r, z))), dm(2.0, da(dm(x, z), dm(r, y)))},
                            {dm(2.0, da(dm(x, y), dm(r, z))), ds(1.0, dm(2.0, da(dm(x, x), dm(z, z)))), dm(2.0, ds(dm(y, z), dm(r, x)))},
                            {dm(2.0, ds(dm(x, z), dm(r, y))), dm(2.0, da(dm(y, z), dm(r, x))), ds(1.0, dm(2.0, da(dm(x, x), dm(y, y))))}};
    double M[3][3];
    for (int k = 0; k < 3; k++) {
        const double sk = dm(exp((double)p.scaling_raw[3 * gj + k]), (double)p.scaling_modifier);
        for (int a = 0; a < 3; a++) M[a][k] = dm(R[a][k], sk);
    }
    double Sg[3][3];
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) Sg[a][b] = da(da(dm(M[a][0], M[b][0]), dm(M[a][1], M[b][1])), dm(M[a][2], M[b][2]));
    double t[3];
    for (int a = 0; a < 3; a++) t[a] = da(da(da(dm(vm[a], mx), dm(vm[4 + a], my)), dm(vm[8 + a], mz)), vm[12 + a]);
    const double fx = dd((double)W, dm(2.0, tfx)), fy = dd((double)H, dm(2.0, tfy));
    const double limx = dm(1.3, tfx), limy = dm(1.3, tfy);
    t[0] = dm(fmin(limx, fmax(-limx, dd(t[0], t[2]))), t[2]);
    t[1] = dm(fmin(limy, fmax(-limy, dd(t[1], t[2]))), t[2]);
    const double tz2 = dm(t[2], t[2]);
    const double Jr[2][3] = {{dd(fx, t[2]), 0.0, -dd(dm(fx, t[0]), tz2)}, {0.0, dd(fy, t[2]), -dd(dm(fy, t[1]), tz2)}};
    // T = Wm @ Jm with Wm = W2C[:3,:3] and Jm rows (Jr[0], Jr[1], 0): T[a][b] = Wm[a][0]*Jr[0][b] + Wm[a][1]*Jr[1][b]
    double T[3][3];
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) T[a][b] = da(dm(vm[4 * 0 + a], Jr[0][b]), dm(vm[4 * 1 + a], Jr[1][b]));
    // cov = T^T Sigma^T T, entries (0,0), (0,1), (1,1)
    double c00 = 0.0, c01 = 0.0, c11 = 0.0;
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) {
        c00 = da(c00, dm(dm(T[a][0], Sg[b][a]), T[b][0]));
        c01 = da(c01, dm(dm(T[a][0], Sg[b][a]), T[b][1]));
        c11 = da(c11, dm(dm(T[a][1], Sg[b][a]), T[b][1]));
    }
    const double cx = da(c00, 0.3), cz = da(c11, 0.3);
    const double det = ds(dm(cx, cz), dm(c01, c01));
    const double mid = dm(0.5, da(cx, cz));
    const double root = __dsqrt_rn(fmax(0.1, ds(dm(mid, mid), det)));
    const float s1 = (float)__dsqrt_rn(da(mid, root)), s2 = (float)__dsqrt_rn(ds(mid, root));   // fp32 like the tensor .item() reads
    // peak at (clamp(int(v)), clamp(int(u))): .long() truncates toward zero (general_utils.py:275-278)
    const float u = p.poses_2d[2 * (size_t)i], vv = p.poses_2d[2 * (size_t)i + 1];
    const int xc = min(max((int)u, 0), W - 1), yc = min(max((int)vv, 0), H - 1);
    const int ry = gauss_radius(s1), rx = gauss_radius(s2);
    const int x0 = max(0, xc - rx), x1 = min(W - 1, xc + rx), y0 = max(0, yc - ry), y1 = min(H - 1, yc + ry);
    roi_rect[4 * (size_t)i] = x0; roi_rect[4 * (size_t)i + 1] = y0; roi_rect[4 * (size_t)i + 2] = x1 - x0 + 1; roi_rect[4 * (size_t)i + 3] = y1 - y0 + 1;
    roi_sigma[2 * (size_t)i] = s1; roi_sigma[2 * (size_t)i + 1] = s2;
    roi_center[2 * (size_t)i] = xc; roi_center[2 * (size_t)i + 1] = yc;
    roi_size[i] = (long long)(x1 - x0 + 1) + (y1 - y0 + 1);      // factored storage: col[h] | row[w]
}

// Response at index idx of scipy's correlate1d(mode='reflect') with the truncated Gaussian (radius taps each side, unnormalised
// weights term[0 .. 2 radius], 1 / their sum) to a unit delta at pos on an axis of length n:
//   sum_k w[k] * [reflect(idx + k - radius) == pos].
__device__ double reflect_response(int idx, int pos, int n, int radius, const double* __restrict__ term, double wsum_inv) {
    // At most three taps can land on pos: through the mirror below 0 (src = -pos - 1), directly (src = pos) and through the mirror
    // above n - 1 (src = 2n - 1 - pos).  Their k = src - idx + radius ascend in that order, which is the order the sequential
    // loop over k (scipy's, and the host specification's) adds them in -- same terms, same order, same bits, O(1) instead of O(radius).
    double acc = 0.0;
    if (2 * radius >= n) {                  // window wider than the axis: a tap can be mirrored twice -- the literal loop (never a SkelSplat shape)
        for (int k = 0; k <= 2 * radius; k++) {
            int src = idx + k - radius;
            if (src < 0) src = -src - 1;
            if (src >= n) src = 2 * n - 1 - src;
            if (src == pos) acc += term[k] * wsum_inv;
        }
        return acc;
    }
    const int k_lo = -pos - 1 - idx + radius, k_mid = pos - idx + radius, k_hi = 2 * n - 1 - pos - idx + radius;
    if (k_lo >= 0 && k_lo <= 2 * radius) acc += term[k_lo] * wsum_inv;        // src = -pos - 1 < 0 always
    if (k_mid >= 0 && k_mid <= 2 * radius) acc += term[k_mid] * wsum_inv;      // src = pos in [0, n)
    if (k_hi >= 0 && k_hi <= 2 * radius) acc += term[k_hi] * wsum_inv;        // src = 2n - 1 - pos >= n always
    return acc;
}

// FACTORED patches.  Inside its window the heatmap is the product of a column profile and a row profile (a separable filter
// applied to a delta, then one scalar normalisation), so a patch is stored as h + w floats  col[h] | row[w]  and the heatmap
// value at window pixel (a, b) is DEFINED as the single fp32 product col[a] * row[b]  (skelsplat_b200/heatmaps.py: same
// definition; it differs from fl32(fl32(col*row) / denom) by <= ~1.5 ulp).  The fused optimiser keeps the profiles of a frame
// in shared memory (~25 KB instead of ~410 KB of patches per frame in HBM).
// One warp per (frame, view, joint) patch: col[a] = fl32(255 * response_y(a)) (scipy stores pass 1 in fp32),
// row[b] = fl32(response_x(b) / (max col * max response_x + 1e-8)) -- per-channel min-max normalisation with min == 0
// (the window never covers the whole image).
constexpr int ROI_WARPS = 4;
__global__ void __launch_bounds__(ROI_WARPS * 32)
roi_fill_kernel(int n_patches, const int* __restrict__ roi_rect, const float* __restrict__ roi_sigma, const int* __restrict__ roi_center,
                const long long* __restrict__ roi_offset, ssb_cameras cams, int V, int J, float* __restrict__ roi_data,
                long long capacity, int* __restrict__ status)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * ROI_WARPS + warp;
    if (i >= n_patches) return;
    const int v = (i / J) % V;
    const int W = cams.dims ? cams.dims[2 * v] : cams.W0, H = cams.dims ? cams.dims[2 * v + 1] : cams.H0;
    const int x0 = roi_rect[4 * (size_t)i], y0 = roi_rect[4 * (size_t)i + 1], w = roi_rect[4 * (size_t)i + 2], h = roi_rect[4 * (size_t)i + 3];
    const double s1 = (double)roi_sigma[2 * (size_t)i], s2 = (double)roi_sigma[2 * (size_t)i + 1];
    const int xc = roi_center[2 * (size_t)i], yc = roi_center[2 * (size_t)i + 1];
    const int ry = (int)(4.0 * s1 + 0.5), rx = (int)(4.0 * s2 + 0.5);
    __shared__ double s_term[ROI_WARPS][2][2 * 128 + 1];
    __shared__ double s_rowd[ROI_WARPS][256];
    if (w > 256 || h > 256 || rx > 128 || ry > 128) {      // sigma > 31 px: not a SkelSplat regime
        if (status && lane == 0) atomicOr(status, (int)SSB_STATUS_ROI_TOO_WIDE);
        return;
    }
    if (capacity >= 0 && roi_offset[i] + (long long)(w + h) > capacity) {       // packed buffer too small: nothing is written
        if (status && lane == 0) atomicOr(status, (int)SSB_STATUS_ROI_OVERFLOW);
        return;
    }
    // normalisation sums of the two truncated kernels: one term per lane, then summed by ONE lane in the sequential order
    // k = -r..r (the order scipy's / the host generator's loop uses), so the value does not depend on the launch shape
    for (int k = lane; k <= 2 * ry; k += 32) s_term[warp][0][k] = exp(-0.5 / (s1 * s1) * (double)((k - ry) * (k - ry)));
    for (int k = lane; k <= 2 * rx; k += 32) s_term[warp][1][k] = exp(-0.5 / (s2 * s2) * (double)((k - rx) * (k - rx)));
    __syncwarp();
    double wsum = 0.0;
    if (lane < 2) {
        const int r = lane ? rx : ry;
        for (int k = 0; k <= 2 * r; k++) wsum += s_term[warp][lane][k];
    }
    const double wy = __shfl_sync(0xFFFFFFFFu, wsum, 0), wx = __shfl_sync(0xFFFFFFFFu, wsum, 1);
    float* dst = roi_data + roi_offset[i];
    float mc = 0.f; double mr = 0.0;
    for (int t = lane; t < h; t += 32) {
        const float c = (float)(255.0 * reflect_response(y0 + t, yc, H, ry, s_term[warp][0], 1.0 / wy));
        dst[t] = c;
        mc = fmaxf(mc, c);
    }
    for (int t = lane; t < w; t += 32) {
        const double r = reflect_response(x0 + t, xc, W, rx, s_term[warp][1], 1.0 / wx);
        s_rowd[warp][t] = r;
        mr = fmax(mr, r);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { mc = fmaxf(mc, __shfl_xor_sync(0xFFFFFFFFu, mc, o)); mr = fmax(mr, __shfl_xor_sync(0xFFFFFFFFu, mr, o)); }
    // max of the rounded products == rounded product of the maxima (all factors >= 0, rounding is monotonic)
    const float denom = (float)((double)mc * mr) + 1e-8f;
    __syncwarp();
    for (int t = lane; t < w; t += 32) dst[h + t] = (float)(s_rowd[warp][t] / (double)denom);
}

// Exclusive scan of the patch sizes into packed offsets (+ the total), one CTA of 32 warps: warp w owns the contiguous
// segment [w*seg, (w+1)*seg) and walks it 32 elements at a time (coalesced), first to get the segment sums, then -- after a
// scan of the 32 segment sums -- to write the offsets.  n = F*V*J is ~1e5: ~10 us, and it keeps the
// detections -> ROIs -> optimiser pipeline free of host round trips.
constexpr int SCAN_THREADS = 1024;
__global__ void __launch_bounds__(SCAN_THREADS)
roi_offsets_kernel(long long n, const long long* __restrict__ size, long long* __restrict__ offset, long long* __restrict__ total)
{
    __shared__ long long s_seg[SCAN_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long seg = ((n + SCAN_THREADS / 32 - 1) / (SCAN_THREADS / 32) + 31) / 32 * 32;      // multiple of 32 elements
    const long long b = seg * warp, e = (b + seg < n) ? b + seg : n;
    long long sum = 0;
    for (long long i = b + lane; i < e; i += 32) sum += size[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
    if (lane == 0) s_seg[warp] = sum;
    __syncthreads();
    if (warp == 0) {                                               // inclusive scan of the segment sums
        long long w = s_seg[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const long long t = __shfl_up_sync(0xFFFFFFFFu, w, o); if (lane >= o) w += t; }
        s_seg[lane] = w;
    }
    __syncthreads();
    long long run = warp ? s_seg[warp - 1] : 0;                    // elements before this segment
    for (long long i0 = b; i0 < e; i0 += 32) {
        const long long i = i0 + lane;
        const long long v = (i < e) ? size[i] : 0;
        long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const long long t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += t; }
        if (i < e) offset[i] = run + incl - v;
        run += __shfl_sync(0xFFFFFFFFu, incl, 31);
    }
    if (threadIdx.x == 0 && total) *total = s_seg[SCAN_THREADS / 32 - 1];
}

}  // namespace ssb

using namespace ssb;

extern "C" {

int ssb_triangulate_dlt(int n_frames, int V, int J, const double* P, const double* poses_2d, double* out_xyz, void* stream_) {
    if (n_frames < 0 || V < 2 || V > DLT_MAXV || J <= 0) return SSB_ERR_INVALID;
    if (n_frames == 0) return SSB_OK;
    if (!P || !poses_2d || !out_xyz) return SSB_ERR_INVALID;
    const int n = n_frames * J;
    dlt_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream_>>>(n_frames, V, J, P, poses_2d, out_xyz);
    return ssb_set_cuda_error(cudaGetLastError());
}

int ssb_heatmap_roi_rects(int n_frames, int J, const ssb_cameras* cams, const float* xyz, const float* scaling_raw,
                          const float* rotation_raw, const float* poses_2d, float scaling_modifier,
                          int* roi_rect, float* roi_sigma, int* roi_center, int64_t* roi_size, void* stream_) {
    if (!cams || n_frames < 0 || J <= 0 || cams->n_views <= 0) return SSB_ERR_INVALID;
    if (n_frames == 0) return SSB_OK;
    if (!xyz || !scaling_raw || !rotation_raw || !poses_2d || !roi_rect || !roi_sigma || !roi_center || !roi_size) return SSB_ERR_INVALID;
    RoiParams p;
    p.F = n_frames; p.V = cams->n_views; p.J = J; p.xyz = xyz; p.scaling_raw = scaling_raw; p.rotation_raw = rotation_raw;
    p.poses_2d = poses_2d; p.cams = *cams; p.scaling_modifier = scaling_modifier;
    const int n = n_frames * cams->n_views * J;
    roi_rect_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream_>>>(p, roi_rect, roi_sigma, roi_center, (long long*)roi_size);
    return ssb_set_cuda_error(cudaGetLastError());
}

int ssb_heatmap_roi_offsets(int64_t n, const int64_t* roi_size, int64_t* roi_offset, int64_t* total, void* stream_) {
    if (n < 0) return SSB_ERR_INVALID;
    if (n > 0 && (!roi_size || !roi_offset)) return SSB_ERR_INVALID;
    roi_offsets_kernel<<<1, SCAN_THREADS, 0, (cudaStream_t)stream_>>>((long long)n, (const long long*)roi_size, (long long*)roi_offset,
                                                                      (long long*)total);
    return ssb_set_cuda_error(cudaGetLastError());
}

int ssb_heatmap_roi_fill(int n_frames, int J, const ssb_cameras* cams, const int* roi_rect, const float* roi_sigma,
                         const int* roi_center, const int64_t* roi_offset, float* roi_data, int64_t capacity, int* status,
                         void* stream_) {
    if (!cams || n_frames < 0 || J <= 0 || cams->n_views <= 0) return SSB_ERR_INVALID;
    if (n_frames == 0) return SSB_OK;
    if (!roi_rect || !roi_sigma || !roi_center || !roi_offset || !roi_data) return SSB_ERR_INVALID;
    const int n = n_frames * cams->n_views * J;
    roi_fill_kernel<<<(n + ROI_WARPS - 1) / ROI_WARPS, ROI_WARPS * 32, 0, (cudaStream_t)stream_>>>(n, roi_rect, roi_sigma, roi_center, (const long long*)roi_offset, *cams,
                                                          cams->n_views, J, roi_data, (long long)capacity, status);
    return ssb_set_cuda_error(cudaGetLastError());
}

}  // extern "C"
