// Fused per-frame optimiser: the whole SkelSplat iteration loop in one persistent kernel, sm_100a.
//
// Replaces, per frame, 500 x (render_* + l2_loss_gaussian + limb consistency + autograd.grad +
// gradient bookkeeping) + 125 x torch.optim.Adam.step  (train.py:130-233, ~50 launches and 5 host
// syncs per iteration in the reference) by ONE kernel launch for any number of frames.
//
// Design (DESIGN.md section "optimiser"):
//   * one CTA owns one frame for all iterations; parameters, Adam moments and the per-view gradient
//     slots live in shared memory; nothing but the GT heatmap ROIs is read from HBM/L2 in the loop;
//   * parameters only change every `accumulation_steps` iterations, so the iterations of one step
//     group are independent: they are processed concurrently as SLOTS "slots" (sequential depth 125
//     instead of 500), each with its own binning state;
//   * binning without a sort: every Gaussian's tiles form a rectangle and there are <= 20 Gaussians, so the
//     position of a (Gaussian, tile) pair in the reference's (tile|depth)-sorted list has a closed form
//     (phase B) => same tile lists, in the same order, as the reference's radix sort;
//   * compositing is never materialised: each warp owns an active tile, walks the row pairs that can
//     contribute, evaluates forward + loss + backward per pixel with the GT read from the ROI patch,
//     keeps per-(tile,Gaussian) raw moment sums in registers and reduces them once per tile with a
//     transposing shuffle butterfly (9 shuffles for 8 values) - no atomics, deterministic.  Tile lists of
//     length 1 and 2 (84 % of the pairs) have hand-specialised functions (tile_one, tile_two);
//   * the one-hot feature structure (Gaussian j renders only into channel j; frozen in the
//     reference: scene/gaussian_model.py:159-166,186) collapses the reference's per-channel
//     backward recurrence into one scalar recurrence per pixel;
//   * Adam (torch.optim.Adam semantics, eps=1e-15) runs in-kernel with host-computed fp64 step sizes.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <type_traits>
#include "common.cuh"
#include "api_internal.h"
#include "adam_form.h"

namespace ssb {

#ifndef SSB_OPT_THREADS
#define SSB_OPT_THREADS 512
#endif
constexpr int OPT_THREADS = SSB_OPT_THREADS;
#ifndef SSB_OPT_MIN_CTAS
#define SSB_OPT_MIN_CTAS 2
#endif
#ifndef SSB_PP_N1           // pixels per lane and loop trip in tile_one (2: two independent dependency chains per lane;
#define SSB_PP_N1 2          // the same for tile lists of length 2 measured slower)
#endif
#ifndef SSB_FAST_MAX         // longest tile list with a fully unrolled register-resident tile function compiled in (longer lists:
#define SSB_FAST_MAX 5       // chunked generic path).  ssb_opt_config::max_unrolled_list selects 4 or 5 at run time (tuning knob).
#endif
#ifndef SSB_ROWCULL          // exact row-band culling as warp-uniform pass-loop bounds (0: all 8 passes; bitwise-identical results).
#define SSB_ROWCULL 1        // An earlier per-(pass, entry) predicate form of the same test measured 8 % SLOWER and was dropped.
#endif
#ifndef SSB_PHASE_TIMING     // developer build: per-phase SM cycles of every CTA's thread 0 summed into g_phase_cycles
#define SSB_PHASE_TIMING 0
#endif
#if SSB_PHASE_TIMING
__device__ unsigned long long g_phase_cycles[8];
__device__ unsigned long long g_list_hist[24];      // tiles by list length (index min(n, 23))
#define SSB_PHASE_MARK(i) { if (tid == 0) { const long long t_now = clock64(); atomicAdd(&g_phase_cycles[i], (unsigned long long)(t_now - t_phase)); t_phase = t_now; } }
#else
#define SSB_PHASE_MARK(i)
#endif
// developer experiments (scripts/gpu_tune_opt.py; never in the product build): upper bounds on what removing the GT loads /
// a cheaper exp could buy.  Results are WRONG with either.
#ifdef SSB_EXP_NOGT
#define SSB_GT_COL(p) (0.25f)
#else
#define SSB_GT_COL(p) (*(p))          // plain (L1-cached) load: the profiles are ~25 KB per frame and re-read every iteration.  The
                                      // read-only path (__ldg, LDG.CONSTANT) measured 4 % SLOWER for the whole kernel
#endif
#ifdef SSB_EXP_FASTEXP
#define SSB_EXPF(x) __expf(x)
#else
#define SSB_EXPF(x) expf(x)
#endif
constexpr int MAXJ = 20;
constexpr int MAXV = 8;
constexpr int MAX_SLOTS = 4;
constexpr int MAX_STEPS = 256;
constexpr int FAST = 4;            // tile-list entries whose gradient sums are held in registers at once
constexpr int NPART = 6;           // dmean2D.x, dmean2D.y, dconic.x, dconic.y, dconic.w, dopacity
constexpr int PSTRIDE = 8;         // per-entry partial record: NPART gradient sums + loss term + mask count (as float)

struct StepTable {                 // host-computed in fp64, rounded to fp32 like torch's scalar->tensor ops
    float neg_step_xyz[MAX_STEPS];
    float neg_step_scaling[MAX_STEPS];
    float neg_step_rotation[MAX_STEPS];
    float neg_step_opacity[MAX_STEPS];
    float bc2_sqrt[MAX_STEPS];
};

struct OptParams {
    ssb_opt_config cfg;
    ssb_cameras cams;
    int n_frames, n_steps;
    float* xyz; float* scaling_raw; float* rotation_raw; float* opacity_raw;
    const int* roi_rect; const int64_t* roi_offset; const float* roi_data;
    float* final_loss; int* status;
    // debug accessor (ssb_optimize_frames_debug; NULL in production): the binning of frame dbg_frame at Adam step dbg_step
    int* dbg; int dbg_frame, dbg_step;
    float one_minus_beta1, one_minus_beta2;    // fp32(1 - beta) with the subtraction in fp64, as torch forms the scalar (fp32(1 - 0.999) != 1.0f - 0.999f)
};

// Reduce V (power of two) values across the warp with V-1+log2(32/V) shuffles.  On return every lane
// holds the full sum of value index  j(lane) = sum over halving steps of (lane & off ? n/2 : 0).
template <int V>
__device__ __forceinline__ float warp_multi_reduce(float (&v)[V], int lane) {
    int off = 16;
#pragma unroll
    for (int n = V; n > 1; n >>= 1) {
        const int half = n >> 1;
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < half; i++) {
            const float send = upper ? v[i] : v[i + half];
            const float keep = upper ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(0xFFFFFFFFu, send, off);
        }
        off >>= 1;
    }
    float x = v[0];
    for (; off > 0; off >>= 1) x += __shfl_xor_sync(0xFFFFFFFFu, x, off);
    return x;
}

// Per-slot splat state (structure of arrays in shared memory).
struct SlotSplats {
    float4 geoA[MAXJ];                     // px, py, opacity, -
    float4 geoB[MAXJ];                     // conic x, y, z, -
    uint32_t depth_bits[MAXJ];
    uint2 rectp[MAXJ];                     // tile rect: x0 | y0 << 16, x1 | y1 << 16 (exclusive); all zero when no tile is touched
    uint16_t tiles[MAXJ], offs[MAXJ];      // tiles touched, inclusive scan
    uint8_t rank[MAXJ];                    // depth rank of each Gaussian in (depth bits, id) order
};

// Per (pixel, Gaussian) backward terms (backward.cu:600-636), accumulated as RAW moment sums: with
//   w = G * dL/dalpha / 2 = G * (err - S) * T      (err = rendered - gt of the Gaussian's own channel, S the recurrence of
//                                                    DESIGN.md 4.1 run on err instead of 2 err)
// a record holds  sum w dx, sum w dy, sum w dx^2, sum w dx dy, sum w dy^2, sum w.  Everything that is constant per Gaussian
// (opacity, conic, the 2 of the MSE derivative, -1/2, the NDC scale, 1/N) is applied ONCE after the per-tile records are
// summed (records_to_grads) instead of once per pixel: 11 instead of ~25 instructions per (pixel, Gaussian).
__device__ __forceinline__ void pair_backward(float (&acc)[PSTRIDE], float dx, float dy, float G, float Tb, float err, float S) {
    const float w = G * ((err - S) * Tb);
    const float wx = w * dx, wy = w * dy;
    acc[0] += wx;
    acc[1] += wy;
    acc[2] = fmaf(wx, dx, acc[2]);
    acc[3] = fmaf(wx, dy, acc[3]);
    acc[4] = fmaf(wy, dy, acc[4]);
    acc[5] += w;
}

// Raw moment sums of one Gaussian -> dL/dmean2D (x, y), dL/dconic (xx, xy, yy), dL/dopacity  (all still to be scaled by 1/N).
__device__ __forceinline__ void records_to_grads(const float (&r)[PSTRIDE], float opac, float conx, float cony, float conz,
                                                 float ddelx_dx, float ddely_dy, float (&out)[NPART]) {
    const float t = 2.f * opac * r[0], u = 2.f * opac * r[1];       // sum dL/dG G dx, sum dL/dG G dy
    out[0] = -(conx * t + cony * u) * ddelx_dx;
    out[1] = -(conz * u + cony * t) * ddely_dy;
    out[2] = -opac * r[2];
    out[3] = -opac * r[3];
    out[4] = -opac * r[4];
    out[5] = 2.f * r[5];
}

// pair_alpha (common.cuh; forward.cu:352-364) with the pass-invariant factors dx, dx*conx, dx*cony supplied by the caller:
// the same operations in the same order, hence the same bits.
__device__ __forceinline__ bool pair_alpha_hoisted(float gpy, float conz, float opac, float dx, float dxcx, float dxcy, float pyf,
                                                   float& dy, float& G, float& alpha) {
    dy = __fsub_rn(gpy, pyf);
    float t = __fmul_rn(dy, __fmul_rn(dy, conz));
    t = __fmaf_rn(dx, dxcx, t);
    const float power = __fmaf_rn(t, -0.5f, -__fmul_rn(dy, dxcy));
    if (power > 0.0f) return false;
    if (power < -5.55f && opac <= 1.0f) return false;
    G = SSB_EXPF(power);
    alpha = fminf(ALPHA_MAX, __fmul_rn(opac, G));
    return !(alpha < ALPHA_MIN);
}

// Loss-mask count, loss term and raw moment sums of one contributing (pixel, Gaussian) pair.  The loss term and the count
// are only ever used as totals over all Gaussians of a view (phase D), so a tile with several Gaussians keeps ONE pair of
// them (in the record of its first entry) and 6 moment sums per entry: 6N + 2 instead of 8N live registers.
__device__ __forceinline__ void pair_accumulate(float (&acc)[PSTRIDE], float& loss, float& count, float dx, float dy, float G, float Tb,
                                                float err, float S, float gt) {
    const float gpos = fmaxf(gt, 0.f);
    count += (gt > 0.f) ? 0.f : 1.f;                               // mask pixel outside {gt > 0} (exact in fp32: < 2^24)
    loss = fmaf(-gpos, gpos, fmaf(err, err, loss));                // err^2 - [gt > 0] gt^2
    pair_backward(acc, dx, dy, G, Tb, err, S);
}

// One reduction per (tile, entry): 8 values (6 gradient sums, loss term, mask count) in 9 shuffles; 8 lanes store the totals.
__device__ __forceinline__ void reduce_store_partial(const float (&acc)[PSTRIDE], float* __restrict__ dst, int lane) {
    float r8[8] = {acc[0], acc[1], acc[2], acc[3], acc[4], acc[5], acc[6], acc[7]};
    const float tot = warp_multi_reduce<8>(r8, lane);
    const int idx = ((lane & 16) ? 4 : 0) | ((lane & 8) ? 2 : 0) | ((lane & 4) ? 1 : 0);
    if ((lane & 3) == 0) dst[idx] = tot;
}

// A tile whose list has exactly N <= FAST Gaussians: everything per-entry lives in registers, loops are fully unrolled.
// The GT patch addressing is hoisted out of the pass loop: per entry a base offset, a row stride for two rows, and the
// range of passes whose row falls inside the patch (columns are pass-invariant for a lane).
template <int N, int PP>
__device__ __forceinline__ void tile_fast(const SlotSplats& sp, const uint16_t* __restrict__ list, const int4* __restrict__ roi_v,
                                          const float* const* __restrict__ fac_v,
                                          int lx, int ly0, int W, int H, float* __restrict__ part_out, int lane)
{
    // The loss term is accumulated on every step although only the last one reports it: a second instantiation without it
    // (one register less per entry) measured 10 % SLOWER overall (more spills in the merged kernel, larger I-cache footprint).
    // PP pixels per lane and loop trip (two consecutive passes when PP == 2): the two pixels are independent dependency
    // chains (ILP), share the per-Gaussian shared-memory loads and accumulate into the same per-entry sums.
    // Row culling: a pass (= two pixel rows of the tile) is run only if it can intersect the alpha >= 1/255 row band
    // [rlo, rhi] of at least one listed Gaussian (phase A; exact, see there).  The test is a warp-uniform LOOP BOUND, not a
    // per-pass predicate -- the per-pass form cost more than it saved.  ~37 % of the (pass, entry) pairs at H36M scale
    // contain no contributing pixel at all.
    int gid[N];
    float grow[N];                                                 // row-profile value of the lane's column in patch u (0 outside it)
    const float* gptr[N];                                          // column profile of patch u at the lane's row of pass 0 (may lie outside it)
    typedef typename std::conditional<(N > 4), unsigned long long, unsigned>::type mask_t;
    mask_t gmask = 0;                                              // bit (8u + pass): the lane's ROW of that pass lies in patch u
    const int ty0 = ly0 - (lane >> 4);                             // tile origin row (warp-uniform)
    int vlo = TILE / 2, vhi = -1;
#pragma unroll
    for (int u = 0; u < N; u++) {
        const int g = list[u];
        gid[u] = g;
        if (SSB_ROWCULL) {
            vlo = min(vlo, (__float_as_int(sp.geoA[g].w) - ty0) >> 1);        // first pass with ty0 + 2 pass + 1 >= rlo
            vhi = max(vhi, (__float_as_int(sp.geoB[g].w) - ty0) >> 1);        // last pass with ty0 + 2 pass <= rhi
        }
        const int4 roi = roi_v[g];
        const int rx = lx - roi.x, ry0 = ly0 - roi.y;
        int plo = ry0 < 0 ? ((1 - ry0) >> 1) : 0;                 // first pass with ry0 + 2*pass >= 0
        int phi = (roi.w - ry0 + 1) >> 1;                          // first pass with ry0 + 2*pass >= h
        phi = phi > TILE / 2 ? TILE / 2 : phi;
        if (phi > plo) gmask |= (mask_t)(((1u << (phi - plo)) - 1u) << plo) << (8 * u);
        const float* fp = fac_v[g];                                // col[h] | row[w]
        grow[u] = ((unsigned)rx < (unsigned)roi.z) ? fp[roi.w + rx] : 0.f;
        gptr[u] = fp + ry0;
        asm volatile("" : "+l"(gptr[u]));   // keep the finished 64-bit pointer (else base + offset is re-derived per load)
    }
    float accv[N][PSTRIDE];           // [u][6], [u][7] (loss term, count) are used for u == 0 only: see pair_accumulate
#pragma unroll
    for (int u = 0; u < N; u++)
#pragma unroll
        for (int q = 0; q < PSTRIDE; q++) accv[u][q] = 0.f;
    if (lx < W) {
        const float pxf = (float)lx;
        if (!SSB_ROWCULL) { vlo = 0; vhi = TILE / 2 - 1; }
        vlo = max(vlo, 0);
        vhi = min(vhi, min(TILE / 2 - 1, (H - 1 - ty0) >> 1));   // and the pass's first row inside the image
        for (int pass0 = vlo; pass0 <= vhi; pass0 += PP) {
            // GT values of the listed Gaussians' channels at the lane's pixels: issued first, consumed only in the backward
            // replay, so the L2 latency hides behind the forward math (a pixel outside a patch reads nothing and gets 0)
            float gtv[PP][N];
            const mask_t gm = gmask >> pass0;
#pragma unroll
            for (int q = 0; q < PP; q++) {
#pragma unroll
                for (int u = 0; u < N; u++) {
                    gtv[q][u] = 0.f;
                    if (gm & ((mask_t)1 << (8 * u + q))) gtv[q][u] = SSB_GT_COL(gptr[u] + 2 * (pass0 + q)) * grow[u];
                }
            }
            float pyf[PP];
            bool live[PP];
#pragma unroll
            for (int q = 0; q < PP; q++) {
                const int py = ly0 + 2 * (pass0 + q);
                pyf[q] = (float)py;
                live[q] = (py < H) && (q == 0 || pass0 + q <= vhi);
            }
            float al[PP][N], Gv[PP][N], Tb[PP][N], T[PP];
            unsigned ok = 0u;                                       // bit (q*N + u)
            bool done[PP];
#pragma unroll
            for (int q = 0; q < PP; q++) { T[q] = 1.0f; done[q] = !live[q]; }
            // forward replay (forward.cu:330-386): alpha, G and the transmittance before each accumulated Gaussian
#pragma unroll
            for (int u = 0; u < N; u++) {
                const float4 A = sp.geoA[gid[u]], B = sp.geoB[gid[u]];
#pragma unroll
                for (int q = 0; q < PP; q++) {
                    al[q][u] = 0.f; Gv[q][u] = 0.f; Tb[q][u] = 0.f;
                    if (!done[q]) {
                        float dx, dy, G, alpha;
                        if (pair_alpha(A.x, A.y, B.x, B.y, B.z, A.z, pxf, pyf[q], dx, dy, G, alpha)) {
                            if (u == 0) {       // T == 1: test_T = 1 - alpha >= 0.01 can never fall below T_EPS
                                al[q][u] = alpha; Gv[q][u] = G; Tb[q][u] = 1.0f; ok |= 1u << (q * N + u); T[q] = __fsub_rn(1.0f, alpha);
                            } else {
                                const float test_T = __fmul_rn(T[q], __fsub_rn(1.0f, alpha));
                                if (test_T < T_EPS) done[q] = true;
                                else { al[q][u] = alpha; Gv[q][u] = G; Tb[q][u] = T[q]; ok |= 1u << (q * N + u); T[q] = test_T; }
                            }
                        }
                    }
                }
            }
            if (ok == 0u) continue;
            // backward replay (backward.cu:536-636) with the one-hot scalar recurrence
            float S[PP], last_alpha[PP], last_g[PP];
#pragma unroll
            for (int q = 0; q < PP; q++) { S[q] = 0.f; last_alpha[q] = 0.f; last_g[q] = 0.f; }
#pragma unroll
            for (int u = N - 1; u >= 0; u--) {
                if (ok & ((1u << u) | (PP == 2 ? (1u << (N + u)) : 0u))) {
                    const float2 mxy = *reinterpret_cast<const float2*>(&sp.geoA[gid[u]]);
                    const float dx = __fsub_rn(mxy.x, pxf);
#pragma unroll
                    for (int q = 0; q < PP; q++) {
                        if ((ok >> (q * N + u)) & 1u) {
                            const float dy = __fsub_rn(mxy.y, pyf[q]);
                            const float gt = gtv[q][u];
                            const float err = fmaf(al[q][u], Tb[q][u], -gt);    // rendered value of channel g minus GT
                            S[q] = fmaf(last_alpha[q], last_g[q] - S[q], S[q]);  // S <- a_last g_last + (1 - a_last) S
                            last_g[q] = err; last_alpha[q] = al[q][u];
                            pair_accumulate(accv[u], accv[0][6], accv[0][7], dx, dy, Gv[q][u], Tb[q][u], err, S[q], gt);
                        }
                    }
                }
            }
        }
    }
#pragma unroll
    for (int u = 0; u < N; u++) reduce_store_partial(accv[u], part_out + (size_t)u * PSTRIDE, lane);
}

// A tile whose list holds ONE Gaussian (46 % of all (tile, Gaussian) entries at H36M scale).  The transmittance before it is 1
// and 1 - alpha >= 0.01 > T_EPS, so nothing of the generic bookkeeping (done flags, T, the recurrence S) exists: forward and
// backward of a pixel collapse into one predicated block -- ~45 instead of ~75 instructions per pixel.  Bit-identical to
// tile_fast<1, PP> (same operations on the same operands, the dropped ones are multiplications by 1 and additions of 0).
template <int PP>
__device__ __forceinline__ void tile_one(const SlotSplats& sp, int g, const int4 roi, const float* __restrict__ fp,
                                         int lx, int ly0, int W, int H, float* __restrict__ part_out, int lane)
{
    const int ty0 = ly0 - (lane >> 4);                             // tile origin row (warp-uniform)
    const float4 A = sp.geoA[g], B = sp.geoB[g];
    const int vlo = SSB_ROWCULL ? max((__float_as_int(A.w) - ty0) >> 1, 0) : 0;
    const int vhi = min(SSB_ROWCULL ? ((__float_as_int(B.w) - ty0) >> 1) : TILE / 2 - 1, min(TILE / 2 - 1, (H - 1 - ty0) >> 1));
    const int rx = lx - roi.x, ry0 = ly0 - roi.y;
    int plo = ry0 < 0 ? ((1 - ry0) >> 1) : 0;                     // first pass with ry0 + 2*pass >= 0
    int phi = (roi.w - ry0 + 1) >> 1;                              // first pass with ry0 + 2*pass >= h
    phi = phi > TILE / 2 ? TILE / 2 : phi;
    unsigned gmask = 0u;                                           // bit pass: the lane's ROW of that pass lies in the patch
    if (phi > plo) gmask = ((1u << (phi - plo)) - 1u) << plo;
    // GT = col[row in patch] * row[column in patch]  (one fp32 product: the definition of the factored heatmap, heatmaps.py)
    const float grow = ((unsigned)rx < (unsigned)roi.z) ? fp[roi.w + rx] : 0.f;      // 0 outside the patch's columns => GT 0
    const float* gp = fp + ry0;                                    // column profile at the lane's row of pass 0; only dereferenced under gmask
    asm volatile("" : "+l"(gp));    // keep the finished 64-bit pointer: otherwise base + offset is re-derived for every load
    float acc[PSTRIDE] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (lx < W) {
        const float dx = __fsub_rn(A.x, (float)lx);
        const float dxcx = __fmul_rn(dx, B.x), dxcy = __fmul_rn(dx, B.y);      // pass-invariant factors of pair_alpha's power
        for (int pass0 = vlo; pass0 <= vhi; pass0 += PP) {
            float gtv[PP];
#pragma unroll
            for (int q = 0; q < PP; q++) {
                gtv[q] = 0.f;
                if ((gmask >> (pass0 + q)) & 1u) gtv[q] = SSB_GT_COL(gp + 2 * (pass0 + q)) * grow;
            }
#pragma unroll
            for (int q = 0; q < PP; q++) {
                const int py = ly0 + 2 * (pass0 + q);
                if (py < H && (q == 0 || pass0 + q <= vhi)) {
                    // pair_alpha (forward.cu:352-364), same operation order
                    const float dy = __fsub_rn(A.y, (float)py);
                    float t = __fmul_rn(dy, __fmul_rn(dy, B.z));
                    t = __fmaf_rn(dx, dxcx, t);
                    const float power = __fmaf_rn(t, -0.5f, -__fmul_rn(dy, dxcy));
                    if (!(power > 0.0f) && !(power < -5.55f && A.z <= 1.0f)) {
                        const float G = SSB_EXPF(power);
                        const float alpha = fminf(ALPHA_MAX, __fmul_rn(A.z, G));
                        if (!(alpha < ALPHA_MIN)) {
                            pair_accumulate(acc, acc[6], acc[7], dx, dy, G, 1.0f, alpha - gtv[q], 0.0f, gtv[q]);   // err = rendered - GT
                        }
                    }
                }
            }
        }
    }
    reduce_store_partial(acc, part_out, lane);
}

// A tile whose list holds TWO Gaussians (38 % of the entries at H36M scale): written out without the generic bookkeeping.
// Front Gaussian 0 sees T = 1 (never terminates the pixel); Gaussian 1 sees T1 = 1 - alpha0 and may hit the T < 1e-4 stop; the
// recurrence reduces to S = alpha1 * err1 for Gaussian 0.  Per-entry constants live in registers for the whole tile.
// Bit-identical to tile_fast<2, 1>.
__device__ __forceinline__ void tile_two(const SlotSplats& sp, const uint16_t* __restrict__ list, const int4* __restrict__ roi_v,
                                         const float* const* __restrict__ fac_v,
                                         int lx, int ly0, int W, int H, float* __restrict__ part_out, int lane)
{
    const int ty0 = ly0 - (lane >> 4);                             // tile origin row (warp-uniform)
    const int g0 = list[0], g1 = list[1];
    const float4 A0 = sp.geoA[g0], B0 = sp.geoB[g0], A1 = sp.geoA[g1], B1 = sp.geoB[g1];
    int vlo = 0, vhi = TILE / 2 - 1;
    if (SSB_ROWCULL) {
        vlo = max(min(__float_as_int(A0.w), __float_as_int(A1.w)) - ty0 >> 1, 0);
        vhi = min(max(__float_as_int(B0.w), __float_as_int(B1.w)) - ty0 >> 1, TILE / 2 - 1);
    }
    vhi = min(vhi, (H - 1 - ty0) >> 1);
    const float* gptr[2];
    float grow[2];
    unsigned gmask = 0u;                                           // bit (8u + pass): the lane's ROW of that pass lies in patch u
#pragma unroll
    for (int u = 0; u < 2; u++) {
        const int g = u ? g1 : g0;
        const int4 roi = roi_v[g];
        const int rx = lx - roi.x, ry0 = ly0 - roi.y;
        int plo = ry0 < 0 ? ((1 - ry0) >> 1) : 0;
        int phi = (roi.w - ry0 + 1) >> 1;
        phi = phi > TILE / 2 ? TILE / 2 : phi;
        if (phi > plo) gmask |= (((1u << (phi - plo)) - 1u) << plo) << (8 * u);
        const float* fp = fac_v[g];
        grow[u] = ((unsigned)rx < (unsigned)roi.z) ? fp[roi.w + rx] : 0.f;
        gptr[u] = fp + ry0;
        asm volatile("" : "+l"(gptr[u]));
    }
    float acc0[PSTRIDE] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, acc1[PSTRIDE] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (lx < W) {
        const float pxf = (float)lx;
        const float dx0 = __fsub_rn(A0.x, pxf), dx1 = __fsub_rn(A1.x, pxf);
        const float dxcx0 = __fmul_rn(dx0, B0.x), dxcy0 = __fmul_rn(dx0, B0.y);
        const float dxcx1 = __fmul_rn(dx1, B1.x), dxcy1 = __fmul_rn(dx1, B1.y);
        for (int pass = vlo; pass <= vhi; pass++) {
            const unsigned gm = gmask >> pass;
            float gt0 = 0.f, gt1 = 0.f;
            if (gm & 1u) gt0 = SSB_GT_COL(gptr[0] + 2 * pass) * grow[0];
            if (gm & 0x100u) gt1 = SSB_GT_COL(gptr[1] + 2 * pass) * grow[1];
            const int py = ly0 + 2 * pass;
            if (py < H) {
                const float pyf = (float)py;
                float dy0, G0, a0, dy1, G1, a1;
                const bool v0 = pair_alpha_hoisted(A0.y, B0.z, A0.z, dx0, dxcx0, dxcy0, pyf, dy0, G0, a0);
                const float T1 = v0 ? __fsub_rn(1.0f, a0) : 1.0f;             // 1 - alpha0 >= 0.01: Gaussian 0 never stops the pixel
                bool v1 = pair_alpha_hoisted(A1.y, B1.z, A1.z, dx1, dxcx1, dxcy1, pyf, dy1, G1, a1);
                if (v1 && __fmul_rn(T1, __fsub_rn(1.0f, a1)) < T_EPS) v1 = false;  // forward.cu:366-371: the pixel is done
                float S = 0.f;
                if (v1) {
                    const float err1 = fmaf(a1, T1, -gt1);
                    pair_accumulate(acc1, acc0[6], acc0[7], dx1, dy1, G1, T1, err1, 0.f, gt1);
                    S = a1 * err1;
                }
                if (v0) pair_accumulate(acc0, acc0[6], acc0[7], dx0, dy0, G0, 1.0f, a0 - gt0, S, gt0);
            }
        }
    }
    reduce_store_partial(acc0, part_out, lane);
    reduce_store_partial(acc1, part_out + PSTRIDE, lane);
}

// NT threads per CTA: 512 with two CTAs per SM, or 1024 with one when the binning state (r_capacity) is too large for two.
// HALVES: the per-(tile,Gaussian) records (32 of the 40 B/pair of binning state) are held for SLOTS/HALVES slots at a time: with
// HALVES == 2 the tile phase runs twice per Adam step (slots 0-1, then 2-3) over the same record storage, which keeps two CTAs
// per SM up to r_capacity 1024 (Panoptic).  Same records, same fixed-order sums => bit-identical results.
template <int SLOTS, int NT, int HALVES>
__global__ void __launch_bounds__(NT, (NT >= 1024 ? 1 : SSB_OPT_MIN_CTAS))
optimize_kernel(const __grid_constant__ OptParams p, const __grid_constant__ StepTable tab)
{
    constexpr int NW = NT / 32;
    constexpr int RSLOTS = SLOTS / HALVES;     // slots whose records are resident at once
    static_assert(SLOTS % HALVES == 0, "HALVES must divide SLOTS");
    const int J = p.cfg.J, V = p.cfg.V, RCAP = p.cfg.r_capacity;
    const int frame = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---------------- shared memory ----------------
    __shared__ float s_xyz[MAXJ * 3], s_scal[MAXJ * 3], s_rot[MAXJ * 4], s_opa[MAXJ];       // raw parameters
    __shared__ float s_m[MAXJ * 11], s_v[MAXJ * 11];                                          // Adam moments
    __shared__ float s_grad[MAXJ * 11];                                                       // grads of the step
    __shared__ float s_accg[MAXV][MAXJ * 3];                                                  // accumulated_grads[V,J,3]
    __shared__ float s_act_scale[MAXJ * 3], s_act_q[MAXJ * 4], s_act_qn[MAXJ], s_act_op[MAXJ], s_cov3d[MAXJ * 6];
    __shared__ float s_view[MAXV][16], s_proj[MAXV][16];
    __shared__ int s_W[MAXV], s_H[MAXV], s_slot_view[MAX_SLOTS];
    __shared__ float s_halfW[MAXV], s_halfH[MAXV];
    __shared__ float s_tfx[MAXV], s_tfy[MAXV], s_fx[MAXV], s_fy[MAXV];
    __shared__ int4 s_roi[MAXV][MAXJ];          // x0, y0, w, h
    __shared__ int s_roi_rel[MAXV][MAXJ];        // float offset of the patch's profiles (col[h] | row[w]) relative to the frame's first patch
    __shared__ long long s_roi_base;
    __shared__ const float* s_fac[MAXV][MAXJ];   // per (view, joint): the patch's profiles col[h] | row[w] in global memory (finished pointers).
                                                 // Staging them in shared memory was built and measured: 1.5 % SLOWER than reading them through
                                                 // L1 (a frame's profiles are ~25 KB and re-read 500 times), bit-identical results -- dropped
    __shared__ int s_ngt[MAXV];            // sum_j #{gt > 0}
    __shared__ float s_sgt2[MAXV];         // sum_j sum gt^2
    __shared__ SlotSplats s_sp[SLOTS];
    __shared__ int s_R[SLOTS], s_nact[SLOTS], s_next[SLOTS];
    __shared__ float s_jcnt[SLOTS][MAXJ], s_jloss[SLOTS][MAXJ];
    __shared__ float s_lsum[SLOTS][NW];
    __shared__ int s_status;
    extern __shared__ __align__(16) unsigned char dsm[];
    // dynamic: partial f32[RSLOTS][RCAP*PSTRIDE] (32 B/entry) whose storage is first used by the sort's 32-bit words of all
    // SLOTS slots (dead once the tile lists exist) | per slot: list u16 | inv_pos u16 | tile u16 | start u16  (8 B/entry)
    float* d_part = reinterpret_cast<float*>(dsm);
    uint16_t* d_list = reinterpret_cast<uint16_t*>(d_part + (size_t)RSLOTS * RCAP * PSTRIDE);
    uint16_t* d_inv = d_list + (size_t)SLOTS * RCAP;
    uint16_t* d_tile = d_inv + (size_t)SLOTS * RCAP;
    uint16_t* d_start = d_tile + (size_t)SLOTS * RCAP;
#define SSB_KEYS32(k) (reinterpret_cast<uint32_t*>(d_part) + (size_t)(k) * RCAP)

    // ---------------- load the frame ----------------
    for (int i = tid; i < J * 3; i += NT) { s_xyz[i] = p.xyz[(size_t)frame * J * 3 + i]; s_scal[i] = p.scaling_raw[(size_t)frame * J * 3 + i]; }
    for (int i = tid; i < J * 4; i += NT) s_rot[i] = p.rotation_raw[(size_t)frame * J * 4 + i];
    for (int i = tid; i < J; i += NT) s_opa[i] = p.opacity_raw[(size_t)frame * J + i];
    for (int i = tid; i < J * 11; i += NT) { s_m[i] = 0.f; s_v[i] = 0.f; s_grad[i] = 0.f; }
    for (int i = tid; i < MAXV * MAXJ * 3; i += NT) (&s_accg[0][0])[i] = 0.f;
    for (int i = tid; i < V * 16; i += NT) { s_view[i / 16][i % 16] = p.cams.viewmatrix[i]; s_proj[i / 16][i % 16] = p.cams.projmatrix[i]; }
    if (tid < V) {
        const int W = p.cams.dims ? p.cams.dims[2 * tid] : p.cams.W0, H = p.cams.dims ? p.cams.dims[2 * tid + 1] : p.cams.H0;
        const float tfx = p.cams.tanfov ? p.cams.tanfov[2 * tid] : p.cams.tanfovx0, tfy = p.cams.tanfov ? p.cams.tanfov[2 * tid + 1] : p.cams.tanfovy0;
        s_W[tid] = W; s_H[tid] = H; s_tfx[tid] = tfx; s_tfy[tid] = tfy;
        s_halfW[tid] = 0.5f * W; s_halfH[tid] = 0.5f * H;
        s_fy[tid] = __fdiv_rn((float)H, __fmul_rn(2.0f, tfy));
        s_fx[tid] = __fdiv_rn((float)W, __fmul_rn(2.0f, tfx));
        s_ngt[tid] = 0; s_sgt2[tid] = 0.f;
    }
    for (int i = tid; i < V * J; i += NT) {
        const int v = i / J, j = i % J;
        const size_t o = ((size_t)frame * V + v) * J + j;
        s_roi[v][j] = make_int4(p.roi_rect[4 * o], p.roi_rect[4 * o + 1], p.roi_rect[4 * o + 2], p.roi_rect[4 * o + 3]);
        s_roi_rel[v][j] = (int)(p.roi_offset[o] - p.roi_offset[(size_t)frame * V * J]);   // patches of a frame are packed together
        s_fac[v][j] = p.roi_data + p.roi_offset[o];
    }
    if (tid == 0) s_roi_base = p.roi_offset[(size_t)frame * V * J];
    if (tid == 0) s_status = 0;
    __syncthreads();
    // GT statistics of the loss mask: N_gt = #{gt>0}, S = sum gt^2 (per view; fixed-order per-warp sums)
    for (int v = 0; v < V; v++) {
        int cnt = 0; float sq = 0.f;
        for (int j = 0; j < J; j++) {
            const int w = s_roi[v][j].z, h = s_roi[v][j].w, n = w * h;
            const float* d = p.roi_data + s_roi_base + s_roi_rel[v][j];       // col[h] | row[w]
            for (int i = tid; i < n; i += NT) {
                const int a = i / w, b = i - a * w;
                const float g = __fmul_rn(__ldg(d + a), __ldg(d + h + b));      // the heatmap value: one fp32 product
                if (g > 0.f) { cnt++; sq = fmaf(g, g, sq); }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o); sq += __shfl_xor_sync(0xFFFFFFFFu, sq, o); }
        if (lane == 0) { s_lsum[0][warp] = sq; atomicAdd(&s_ngt[v], cnt); }
        __syncthreads();
        if (tid == 0) { float a = 0.f; for (int w = 0; w < NW; w++) a += s_lsum[0][w]; s_sgt2[v] = a; }
        __syncthreads();
    }

    const int acc = p.cfg.accumulation_steps;     // == SLOTS (host guarantees)
    float last_loss = 0.f;
#if SSB_PHASE_TIMING
    long long t_phase = clock64();
#endif
    for (int step = 0; step < p.n_steps; step++) {
        SSB_PHASE_MARK(7)
        // ============ phase A: activations + projection of every (slot, joint) ============
        if (tid < J) {
            const int j = tid;
            // GaussianModel getters (scene/gaussian_model.py:102-143): exp / normalize / sigmoid
            const float sx = expf(s_scal[3 * j]), sy = expf(s_scal[3 * j + 1]), sz = expf(s_scal[3 * j + 2]);
            const float q0 = s_rot[4 * j], q1 = s_rot[4 * j + 1], q2 = s_rot[4 * j + 2], q3 = s_rot[4 * j + 3];
            const float qn = fmaxf(sqrtf(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3), 1e-12f);   // F.normalize eps
            const float r = q0 / qn, x = q1 / qn, y = q2 / qn, z = q3 / qn;
            s_act_scale[3 * j] = sx; s_act_scale[3 * j + 1] = sy; s_act_scale[3 * j + 2] = sz;
            s_act_q[4 * j] = r; s_act_q[4 * j + 1] = x; s_act_q[4 * j + 2] = y; s_act_q[4 * j + 3] = z;
            s_act_qn[j] = qn;
            s_act_op[j] = 1.0f / (1.0f + expf(-s_opa[j]));
            float cov[6];
            cov3d_from_scale_rot(sx, sy, sz, 1.0f, r, x, y, z, cov);
#pragma unroll
            for (int k = 0; k < 6; k++) s_cov3d[6 * j + k] = cov[k];
        }
        if (tid < SLOTS) s_slot_view[tid] = (step * acc + tid) % V;
        if (tid < SLOTS) s_next[tid] = 0;
        __syncthreads();
        if (tid < SLOTS * J) {
            const int k = tid / J, j = tid % J;
            const int v = (step * acc + k) % V;
            const Splat s = project_gaussian(s_xyz[3 * j], s_xyz[3 * j + 1], s_xyz[3 * j + 2], &s_cov3d[6 * j], s_act_op[j],
                                             s_view[v], s_proj[v], s_W[v], s_H[v], s_tfx[v], s_tfy[v], s_fx[v], s_fy[v],
                                             p.cfg.antialiasing != 0);
            SlotSplats& sp = s_sp[k];
            // Row band outside which the Gaussian cannot reach alpha >= 1/255 (so contributes nothing, forward.cu:358-363):
            // max over dx of power(dx, dy) = -0.5 dy^2 det/conx must be >= -5.55 (pair_alpha's exact early-out; opacity <= 1)
            // => |dy| <= sqrt(11.1 conx/det), widened by 1 % + 1 px against fp32 rounding of the conic.  A near-singular
            // conic (det lost to cancellation) gets no band.
            const float cdet = s.conx * s.conz - s.cony * s.cony;
            const float ey = (s.conx > 0.f && cdet > 1e-4f * s.conx * s.conz) ? 1.01f * sqrtf(11.1f * s.conx / cdet) + 1.0f : 1e9f;
            const int rlo = (int)fminf(fmaxf(ceilf(s.py - ey), 0.0f), 65535.0f), rhi = (int)fminf(fmaxf(floorf(s.py + ey), -1.0f), 65535.0f);
            sp.geoA[j] = make_float4(s.px, s.py, s.opac, __int_as_float(rlo));
            sp.geoB[j] = make_float4(s.conx, s.cony, s.conz, __int_as_float(rhi));
            sp.depth_bits[j] = __float_as_uint(s.depth);
            sp.rectp[j] = s.tiles > 0 ? make_uint2((uint32_t)s.rect.x0 | ((uint32_t)s.rect.y0 << 16), (uint32_t)s.rect.x1 | ((uint32_t)s.rect.y1 << 16))
                                      : make_uint2(0u, 0u);
            sp.tiles[j] = (uint16_t)min(s.tiles, 65535u);
        }
        __syncthreads();
        SSB_PHASE_MARK(0)
        // ============ phase B: binning per slot.  The reference sorts (tile | depth) keys (rasterizer_impl.cu:303-311); with at
        // most MAXJ Gaussians whose tiles form rectangles, the sorted position of an entry has a closed form:
        //   pos(j, tile t) = sum over j' of #{tiles of rect(j') that precede t in row-major order}
        //                  + #{j' : t in rect(j') and (depth, id)(j') < (depth, id)(j)}
        // i.e. the order of the stable sort, without sorting: no compare-exchange network, no barriers (the bitonic sort this
        // replaces was 36 barrier-separated stages = 6 % of the kernel).  Entries are handled one per thread. ============
        if (tid < SLOTS) {
            SlotSplats& sp = s_sp[tid];
            uint32_t a = 0;
            for (int j = 0; j < J; j++) { a += sp.tiles[j]; sp.offs[j] = (uint16_t)min(a, 65535u); }
            // capacity overflow: flag the frame (the host re-runs it with a larger r_capacity) and skip the slot
            if (a > (uint32_t)RCAP) { s_status |= (int)SSB_STATUS_R_OVERFLOW; a = 0; }
            s_R[tid] = (int)a;
        }
        if (tid >= 32 && tid < 32 + SLOTS * J) {      // depth rank: position of (depth bits, id) among the slot's Gaussians
            const int k = (tid - 32) / J, j = (tid - 32) % J;
            SlotSplats& sp = s_sp[k];
            const uint32_t dj = sp.depth_bits[j];
            int r = 0;
            for (int o = 0; o < J; o++) { const uint32_t d = sp.depth_bits[o]; r += (d < dj || (d == dj && o < j)) ? 1 : 0; }
            sp.rank[j] = (uint8_t)r;
        }
        __syncthreads();
        int nsort = 32, lgsort = 5;
        {
            int Rmax = 0;
#pragma unroll
            for (int k = 0; k < SLOTS; k++) Rmax = max(Rmax, s_R[k]);
            while (nsort < Rmax) { nsort <<= 1; lgsort++; }       // power of two >= every slot's R: thread -> (slot, entry) split
        }
        for (int i = tid; i < SLOTS * nsort; i += NT) {
            const int k = i >> lgsort, idx = i & (nsort - 1);      // idx: emission index (Gaussian-major, row-major in its rect)
            if (idx < s_R[k]) {
                const SlotSplats& sp = s_sp[k];
                int j = 0;
                while ((int)sp.offs[j] <= idx) j++;
                const int local = idx - (j ? (int)sp.offs[j - 1] : 0);
                const uint2 rj = sp.rectp[j];
                const int x0 = (int)(rj.x & 0xFFFFu), y0 = (int)(rj.x >> 16), wj = (int)(rj.y & 0xFFFFu) - x0;
                const int yy = local / wj;
                const int y = y0 + yy, x = x0 + (local - yy * wj);
                const int rank_j = sp.rank[j];
                int pos = 0;
                for (int o = 0; o < J; o++) {
                    const uint2 ro = sp.rectp[o];                  // empty (0,0,0,0) for a Gaussian that touches no tile
                    const int ox0 = (int)(ro.x & 0xFFFFu), oy0 = (int)(ro.x >> 16), ox1 = (int)(ro.y & 0xFFFFu), oy1 = (int)(ro.y >> 16);
                    const int ow = ox1 - ox0;
                    pos += min(max(y - oy0, 0), oy1 - oy0) * ow;               // whole rows above y
                    if (y >= oy0 && y < oy1) {
                        pos += min(max(x - ox0, 0), ow);                       // same row, left of x
                        if (x >= ox0 && x < ox1 && (int)sp.rank[o] < rank_j) pos++;   // same tile, nearer Gaussian
                    }
                }
                SSB_KEYS32(k)[pos] = ((uint32_t)y << 8) | (uint32_t)x;         // packed tile coordinate of sorted entry pos
                d_list[(size_t)k * RCAP + pos] = (uint16_t)j;
                d_inv[(size_t)k * RCAP + idx] = (uint16_t)pos;
            }
        }
        __syncthreads();
        SSB_PHASE_MARK(1)
        if (warp < SLOTS) {      // warp k: ordered compaction of the tile runs of slot k
            const int k = warp, R = s_R[k];
            const uint32_t* K = SSB_KEYS32(k);
            int nact = 0;
            for (int base = 0; base < R; base += 32) {
                const int i = base + lane;
                bool start = false; uint32_t tile = 0;
                if (i < R) { tile = K[i]; start = (i == 0) || (K[i - 1] != tile); }
                const uint32_t m = __ballot_sync(0xFFFFFFFFu, start);
                if (start) {
                    const int a = nact + __popc(m & ((1u << lane) - 1u));
                    d_tile[(size_t)k * RCAP + a] = (uint16_t)tile;
                    d_start[(size_t)k * RCAP + a] = (uint16_t)i;
                }
                nact += __popc(m);
            }
            if (lane == 0) s_nact[k] = nact;
        }
        __syncthreads();
        if (p.dbg && frame == p.dbg_frame && step == p.dbg_step) {
            // per slot k, 4 + 4 RCAP int32: (R, active tiles, view, status) | list[RCAP] | inv_pos[RCAP] | tile[RCAP] | start[RCAP]
            for (int i = tid; i < SLOTS * RCAP; i += NT) {
                const int k = i / RCAP, e = i - k * RCAP;
                int* o = p.dbg + (size_t)k * (4 + 4 * RCAP) + 4;
                o[e] = e < s_R[k] ? (int)d_list[i] : -1;
                o[RCAP + e] = e < s_R[k] ? (int)d_inv[i] : -1;
                o[2 * RCAP + e] = e < s_nact[k] ? (int)d_tile[i] : -1;
                o[3 * RCAP + e] = e < s_nact[k] ? (int)d_start[i] : -1;
            }
            if (tid < SLOTS) {
                int* o = p.dbg + (size_t)tid * (4 + 4 * RCAP);
                o[0] = s_R[tid]; o[1] = s_nact[tid]; o[2] = s_slot_view[tid]; o[3] = s_status;
            }
        }

        SSB_PHASE_MARK(3)
        // ============ phase C: tiles.  One warp per active tile, handed out dynamically (tile lists differ in length);
        // every result is a per-(tile,Gaussian) record, so the schedule does not influence any sum ============
        float s8[PSTRIDE] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
        for (int half = 0; half < HALVES; half++) {
        {
            const bool unroll5 = p.cfg.max_unrolled_list != 4;          // 0 (default) or 5: lists of five take tile_fast<5>
            // one counter per slot; a warp starts on slot (warp mod SLOTS) and moves on when that slot's tiles are handed out,
            // so everything that depends on the slot only (view, image size, splat table) is loaded once per slot, not per tile
            for (int kk = 0; kk < RSLOTS; kk++) {
            const int k = half * RSLOTS + ((warp + kk) & (RSLOTS - 1));
            const int nact = s_nact[k];
            const int v = s_slot_view[k];
            const int W = s_W[v], H = s_H[v];
            const SlotSplats& sp = s_sp[k];
            for (;;) {
                int a = 0;
                if (lane == 0) a = atomicAdd(&s_next[k], 1);
                a = __shfl_sync(0xFFFFFFFFu, a, 0);
                if (a >= nact) break;
                const int tile = d_tile[(size_t)k * RCAP + a];                 // packed (ty << 8) | tx
                const int e0 = d_start[(size_t)k * RCAP + a];
                const int e1 = (a + 1 < nact) ? (int)d_start[(size_t)k * RCAP + a + 1] : s_R[k];
                const int n = e1 - e0;
                const uint16_t* list = d_list + (size_t)k * RCAP + e0;
#if SSB_PHASE_TIMING
                if (lane == 0) atomicAdd(&g_list_hist[n < 23 ? n : 23], 1ull);
#endif
                const int lx = (tile & 255) * TILE + (lane & 15), ly0 = (tile >> 8) * TILE + (lane >> 4);
                float* part_out = d_part + ((size_t)(k - half * RSLOTS) * RCAP + e0) * PSTRIDE;
#define SSB_TILE_FAST(NN, PPP) tile_fast<NN, PPP>(sp, list, s_roi[v], s_fac[v], lx, ly0, W, H, part_out, lane);
                if (n == 1) { const int g1 = list[0]; tile_one<SSB_PP_N1>(sp, g1, s_roi[v][g1], s_fac[v][g1], lx, ly0, W, H, part_out, lane); }
                else if (n == 2) tile_two(sp, list, s_roi[v], s_fac[v], lx, ly0, W, H, part_out, lane);
                else if (n == 3 && SSB_FAST_MAX >= 3) SSB_TILE_FAST(3, 1)
                else if (n == 4 && SSB_FAST_MAX >= 4) SSB_TILE_FAST(4, 1)
                else if (n == 5 && SSB_FAST_MAX >= 5 && unroll5) SSB_TILE_FAST(5, 1)
                else if (n == 6 && SSB_FAST_MAX >= 6) SSB_TILE_FAST(6, 1)
#undef SSB_TILE_FAST
                else {
                    // ---------- generic path (long tile lists): entries in chunks of FAST, replayed per chunk
                    for (int c0 = 0; c0 < n; c0 += FAST) {
                        float accv[FAST][PSTRIDE];
#pragma unroll
                        for (int u = 0; u < FAST; u++)
#pragma unroll
                            for (int q = 0; q < PSTRIDE; q++) accv[u][q] = 0.f;
                        for (int pass = 0; pass < TILE / 2; pass++) {
                            const int px = lx, py = ly0 + 2 * pass;
                            if (!(px < W && py < H)) continue;          // no warp-collective operation inside the pass loop
                            const float pxf = (float)px, pyf = (float)py;
                            // forward replay: final transmittance and last contributor (forward.cu:330-386)
                            float T = 1.0f;
                            int last_contributor = 0;
                            for (int e = 0; e < n; e++) {
                                const int g = list[e];
                                const float4 A = sp.geoA[g], B = sp.geoB[g];
                                float dx, dy, G, alpha;
                                if (!pair_alpha(A.x, A.y, B.x, B.y, B.z, A.z, pxf, pyf, dx, dy, G, alpha)) continue;
                                const float test_T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
                                if (test_T < T_EPS) break;
                                T = test_T;
                                last_contributor = e + 1;
                            }
                            // backward replay (backward.cu:536-636) with the one-hot scalar recurrence
                            float S = 0.f, last_alpha = 0.f, last_g = 0.f;
                            for (int e = last_contributor - 1; e >= 0; e--) {
                                const int g = list[e];
                                const float4 A = sp.geoA[g], B = sp.geoB[g];
                                float dx, dy, G, alpha;
                                if (!pair_alpha(A.x, A.y, B.x, B.y, B.z, A.z, pxf, pyf, dx, dy, G, alpha)) continue;
                                T = T / (1.f - alpha);
                                const int4 roi = s_roi[v][g];
                                const int rx = px - roi.x, ry = py - roi.y;
                                float gt = 0.f;
                                if ((unsigned)rx < (unsigned)roi.z && (unsigned)ry < (unsigned)roi.w)
                                    gt = __fmul_rn(s_fac[v][g][ry], s_fac[v][g][roi.w + rx]);
                                const float err = fmaf(alpha, T, -gt);
                                S = fmaf(last_alpha, last_g - S, S);
                                last_g = err; last_alpha = alpha;
                                const int u = e - c0;
                                if (u >= 0 && u < FAST) {
                                    float w[PSTRIDE] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                                    const float gpos = fmaxf(gt, 0.f);
                                    w[7] = (gt > 0.f) ? 0.f : 1.f;
                                    w[6] = fmaf(-gpos, gpos, err * err);
                                    pair_backward(w, dx, dy, G, T, err, S);
#pragma unroll
                                    for (int uu = 0; uu < FAST; uu++)
                                        if (u == uu) {
#pragma unroll
                                            for (int q = 0; q < PSTRIDE; q++) accv[uu][q] += w[q];
                                        }
                                }
                            }
                        }
#pragma unroll
                        for (int u = 0; u < FAST; u++)
                            if (c0 + u < n) reduce_store_partial(accv[u], part_out + (size_t)(c0 + u) * PSTRIDE, lane);   // warp-uniform
                    }
                }
            }
            }
        }
        __syncthreads();

        SSB_PHASE_MARK(4)
        // ============ phase D: per-Gaussian backward chain + gradient bookkeeping ============
        // D1: fixed-order (emission order) sum of each Gaussian's per-tile records
        if (tid < SLOTS * J && tid / J / RSLOTS == half) {
            const int k = tid / J, j = tid % J;
            const SlotSplats& sp = s_sp[k];
            if (sp.tiles[j] > 0) {
                const int o0 = (j == 0) ? 0 : sp.offs[j - 1], o1 = min((int)sp.offs[j], s_R[k]);
                for (int o = o0; o < o1; o++) {
                    const float* pp = d_part + ((size_t)(k - half * RSLOTS) * RCAP + d_inv[(size_t)k * RCAP + o]) * PSTRIDE;
#pragma unroll
                    for (int q = 0; q < PSTRIDE; q++) s8[q] += pp[q];
                }
            }
            s_jloss[k][j] = s8[6];
            s_jcnt[k][j] = s8[7];
        }
        __syncthreads();
        }   // half
        // D2: mask size N of the slot's view, then the chain
        if (tid < SLOTS * J) {
            const int k = tid / J, j = tid % J;
            const int v = (step * acc + k) % V;
            const SlotSplats& sp = s_sp[k];
            float extra = 0.f;
            for (int o = 0; o < J; o++) extra += s_jcnt[k][o];
            const float invN = 1.0f / ((float)s_ngt[v] + extra);      // mean over the loss mask (N < 2^24: exact)
            float gm[3] = {0.f, 0.f, 0.f}, gs[3] = {0.f, 0.f, 0.f}, gq[4] = {0.f, 0.f, 0.f, 0.f}, gop = 0.f;
            if (sp.tiles[j] > 0) {
                float s6[NPART];
                const float4 gA = sp.geoA[j], gB = sp.geoB[j];
                records_to_grads(s8, gA.z, gB.x, gB.y, gB.z, s_halfW[v], s_halfH[v], s6);
#pragma unroll
                for (int q = 0; q < NPART; q++) s6[q] *= invN;
                const SplatGrad sg = gaussian_backward(
                    s_xyz[3 * j], s_xyz[3 * j + 1], s_xyz[3 * j + 2], &s_cov3d[6 * j], true,
                    s_act_scale[3 * j], s_act_scale[3 * j + 1], s_act_scale[3 * j + 2], 1.0f,
                    s_act_q[4 * j], s_act_q[4 * j + 1], s_act_q[4 * j + 2], s_act_q[4 * j + 3],
                    s_view[v], s_proj[v], s_fx[v], s_fy[v], s_tfx[v], s_tfy[v],
                    s6[0], s6[1], s6[2], s6[3], s6[4], 0.f, true);
                gm[0] = sg.dmean[0]; gm[1] = sg.dmean[1]; gm[2] = sg.dmean[2];
                // exp backward
                gs[0] = sg.dscale[0] * s_act_scale[3 * j]; gs[1] = sg.dscale[1] * s_act_scale[3 * j + 1]; gs[2] = sg.dscale[2] * s_act_scale[3 * j + 2];
                // F.normalize backward: (g - y (y.g)) / |x|
                const float dotq = sg.drot[0] * s_act_q[4 * j] + sg.drot[1] * s_act_q[4 * j + 1] + sg.drot[2] * s_act_q[4 * j + 2] + sg.drot[3] * s_act_q[4 * j + 3];
#pragma unroll
                for (int c = 0; c < 4; c++) gq[c] = (sg.drot[c] - s_act_q[4 * j + c] * dotq) / s_act_qn[j];
                // sigmoid backward
                gop = s6[5] * s_act_op[j] * (1.f - s_act_op[j]);
            }
            // accumulated_grads[view] = raster gradient + lambda * limb-consistency gradient (added below)
            s_accg[v][3 * j] = gm[0]; s_accg[v][3 * j + 1] = gm[1]; s_accg[v][3 * j + 2] = gm[2];
            if (k == SLOTS - 1) {   // scaling / rotation / opacity grads: the LAST view of the group only (train.py:177-179)
                s_grad[3 * J + 3 * j] = gs[0]; s_grad[3 * J + 3 * j + 1] = gs[1]; s_grad[3 * J + 3 * j + 2] = gs[2];
#pragma unroll
                for (int c = 0; c < 4; c++) s_grad[6 * J + 4 * j + c] = gq[c];
                s_grad[10 * J + j] = gop;
            }
        }
        __syncthreads();
        // limb-consistency term: same xyz for every slot of the group => same gradient added to each slot's view
        if (tid < SLOTS) {
            const int k = tid, v = (step * acc + k) % V;
            float cons = 0.f;
            const bool last_of_view = (k + V >= SLOTS);      // V < SLOTS: several slots share a view, the last one wins
            if (p.cfg.lambda_consistency != 0.f) {
                for (int h = 0; h < 2; h++) {
                    const int a0 = p.cfg.limb_pairs[4 * h], a1 = p.cfg.limb_pairs[4 * h + 1], b0 = p.cfg.limb_pairs[4 * h + 2], b1 = p.cfg.limb_pairs[4 * h + 3];
                    float da[3], db[3];
#pragma unroll
                    for (int c = 0; c < 3; c++) { da[c] = s_xyz[3 * a0 + c] - s_xyz[3 * a1 + c]; db[c] = s_xyz[3 * b0 + c] - s_xyz[3 * b1 + c]; }
                    const float la = sqrtf(da[0] * da[0] + da[1] * da[1] + da[2] * da[2]);
                    const float lb = sqrtf(db[0] * db[0] + db[1] * db[1] + db[2] * db[2]);
                    const float diff = la - lb;
                    cons += fabsf(diff);
                    const float sgn = (diff > 0.f) ? 1.f : ((diff < 0.f) ? -1.f : 0.f);
                    const float lam = p.cfg.lambda_consistency;
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        const float ga = (la > 0.f) ? lam * sgn * da[c] / la : 0.f;
                        const float gb = (lb > 0.f) ? -lam * sgn * db[c] / lb : 0.f;
                        if (last_of_view) {
                            s_accg[v][3 * a0 + c] += ga; s_accg[v][3 * a1 + c] -= ga;
                            s_accg[v][3 * b0 + c] += gb; s_accg[v][3 * b1 + c] -= gb;
                        }
                    }
                }
            }
            if (k == SLOTS - 1) {
                float a = 0.f, extra = 0.f;
                for (int o = 0; o < J; o++) { a += s_jloss[k][o]; extra += s_jcnt[k][o]; }
                last_loss = (a + s_sgt2[v]) / ((float)s_ngt[v] + extra) + p.cfg.lambda_consistency * cons;
                if (p.final_loss && step == p.n_steps - 1) p.final_loss[frame] = last_loss;
            }
        }
        __syncthreads();
        SSB_PHASE_MARK(5)
        // ============ phase E: Adam (torch.optim.Adam, foreach path; train.py:215-222) ============
        if (tid < J * 11) {
            const int i = tid;
            float g, neg_step;
            float* param;
            if (i < 3 * J) {                 // xyz: mean over the V slots (stale / zero slots included)
                float a = 0.f;
                for (int v = 0; v < V; v++) a += s_accg[v][i];
                g = a / (float)V;
                neg_step = tab.neg_step_xyz[step]; param = &s_xyz[i];
            } else if (i < 6 * J) { g = s_grad[i]; neg_step = tab.neg_step_scaling[step]; param = &s_scal[i - 3 * J]; }
            else if (i < 10 * J) { g = s_grad[i]; neg_step = tab.neg_step_rotation[step]; param = &s_rot[i - 6 * J]; }
            else { g = s_grad[i]; neg_step = tab.neg_step_opacity[step]; param = &s_opa[i - 10 * J]; }
            const float m = fmaf(p.one_minus_beta1, g - s_m[i], s_m[i]);                  // lerp_(grad, 1-beta1)
            const float vv = SSB_ADAM_SECOND_MOMENT(p.one_minus_beta2, g, s_v[i] * p.cfg.beta2);   // mul_(beta2).addcmul_(g, g, 1-beta2)
            s_m[i] = m; s_v[i] = vv;
            const float denom = sqrtf(vv) / tab.bc2_sqrt[step] + p.cfg.eps;
            *param = fmaf(neg_step, m / denom, *param);                                   // addcdiv_(m, denom, -step_size)
        }
        __syncthreads();
    }

    SSB_PHASE_MARK(6)
    // ---------------- write back ----------------
    for (int i = tid; i < J * 3; i += NT) { p.xyz[(size_t)frame * J * 3 + i] = s_xyz[i]; p.scaling_raw[(size_t)frame * J * 3 + i] = s_scal[i]; }
    for (int i = tid; i < J * 4; i += NT) p.rotation_raw[(size_t)frame * J * 4 + i] = s_rot[i];
    for (int i = tid; i < J; i += NT) p.opacity_raw[(size_t)frame * J + i] = s_opa[i];
    if (tid == 0 && p.status) p.status[frame] = s_status;
    (void)last_loss;
}

static size_t opt_dyn_smem(int slots, int rcap, int halves) {
    return (size_t)(slots / halves) * rcap * (4 * PSTRIDE) + (size_t)slots * rcap * (2 * 4);
}
constexpr int OPT_STATIC_SMEM = 19 * 1024;       // static shared memory (17.0 KB) + the 1 KB the driver reserves per CTA, rounded up

}  // namespace ssb

using namespace ssb;

extern "C" {

#if SSB_PHASE_TIMING
// developer build only: cycles per phase (A, B placement, -, B compaction, C tiles, D chain, tail, E Adam)
int ssb_debug_phase_cycles(unsigned long long* out8, int reset) {
    if (cudaMemcpyFromSymbol(out8, g_phase_cycles, sizeof(unsigned long long) * 8) != cudaSuccess) return SSB_ERR_CUDA;
    if (reset) { unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0}; cudaMemcpyToSymbol(g_phase_cycles, z, sizeof(z)); }
    return SSB_OK;
}
int ssb_debug_list_hist(unsigned long long* out24) {
    return cudaMemcpyFromSymbol(out24, g_list_hist, sizeof(unsigned long long) * 24) == cudaSuccess ? SSB_OK : SSB_ERR_CUDA;
}
#endif

size_t ssb_optimize_workspace_bytes(const ssb_opt_config* cfg, int n_frames) {
    (void)cfg;
    return (size_t)(n_frames > 0 ? n_frames : 1) * sizeof(int);     // per-frame status words
}

static int optimize_frames_impl(const ssb_opt_config* cfg, int n_frames, const ssb_cameras* cams, const double* lr_xyz_host,
                                float* xyz, float* scaling_raw, float* rotation_raw, float* opacity_raw,
                                const int* roi_rect, const int64_t* roi_offset, const float* roi_data,
                                float* final_loss, void* workspace, int dbg_frame, int dbg_step, int* dbg_out, void* stream_)
{
    if (!cfg || !cams || !lr_xyz_host || n_frames < 0) return SSB_ERR_INVALID;
    if (cfg->J <= 0 || cfg->J > MAXJ || cfg->V <= 0 || cfg->V > MAXV || cams->n_views != cfg->V) return SSB_ERR_UNSUPPORTED;
    // tile coordinates are packed as (y << 8) | x in 16 bits: at most 256 tiles (4096 px) per axis.  W0/H0 are the maxima
    // over the views (ssb_cameras), so dims[] on the device may not exceed them.
    if (cams->W0 <= 0 || cams->H0 <= 0 || cams->W0 > 4096 || cams->H0 > 4096) return SSB_ERR_UNSUPPORTED;
    if (cfg->accumulation_steps < 1 || cfg->accumulation_steps > MAX_SLOTS || cfg->accumulation_steps == 3) return SSB_ERR_UNSUPPORTED;
    if (cfg->r_capacity < 32 || cfg->r_capacity > 1024 || (cfg->r_capacity % 32)) return SSB_ERR_CAPACITY;
    if (cfg->max_unrolled_list != 0 && cfg->max_unrolled_list != 4 && cfg->max_unrolled_list != 5) return SSB_ERR_INVALID;
    const int n_steps = cfg->iterations / cfg->accumulation_steps;   // trailing iterations never reach an optimiser step
    if (n_steps > MAX_STEPS) return SSB_ERR_CAPACITY;
    for (int i = 0; i < 8; i++) if (cfg->limb_pairs[i] < 0 || cfg->limb_pairs[i] >= cfg->J) return SSB_ERR_INVALID;
    if (n_frames == 0) return SSB_OK;
    if (!xyz || !scaling_raw || !rotation_raw || !opacity_raw || !roi_rect || !roi_offset || !roi_data || !workspace) return SSB_ERR_INVALID;

    // Host scalars exactly as torch.optim.Adam's foreach path computes them (python floats = fp64),
    // then rounded to fp32 where torch hands them to an fp32 tensor op.  The config carries betas and learning rates as fp32;
    // torch sees the decimal literal of the yaml as a python float (0.9, not 0.9f = 0.89999998), so each is taken back to the
    // shortest decimal that round-trips the fp32 value before it enters the fp64 arithmetic.
    auto as_written = [](float x) { char buf[32]; std::snprintf(buf, sizeof(buf), "%.7g", (double)x); return std::strtod(buf, nullptr); };
    const double b1 = as_written(cfg->beta1), b2 = as_written(cfg->beta2);
    const double lr_s = as_written(cfg->lr_scaling), lr_r = as_written(cfg->lr_rotation), lr_o = as_written(cfg->lr_opacity);
    StepTable tab;
    for (int s = 0; s < n_steps; s++) {
        const double step = (double)(s + 1);
        const double bc1 = 1.0 - std::pow(b1, step);
        const double bc2 = 1.0 - std::pow(b2, step);
        const int it = (s + 1) * cfg->accumulation_steps;          // lr is taken at the stepping iteration
        tab.neg_step_xyz[s] = (float)(-(lr_xyz_host[it] / bc1));
        tab.neg_step_scaling[s] = (float)(-(lr_s / bc1));
        tab.neg_step_rotation[s] = (float)(-(lr_r / bc1));
        tab.neg_step_opacity[s] = (float)(-(lr_o / bc1));
        tab.bc2_sqrt[s] = (float)std::sqrt(bc2);
    }
    OptParams p;
    p.cfg = *cfg; p.cams = *cams; p.n_frames = n_frames; p.n_steps = n_steps;
    p.xyz = xyz; p.scaling_raw = scaling_raw; p.rotation_raw = rotation_raw; p.opacity_raw = opacity_raw;
    p.roi_rect = roi_rect; p.roi_offset = roi_offset; p.roi_data = roi_data; p.final_loss = final_loss;
    p.status = (int*)workspace;
    p.dbg = dbg_out; p.dbg_frame = dbg_frame; p.dbg_step = dbg_step;
    p.one_minus_beta1 = (float)(1.0 - b1); p.one_minus_beta2 = (float)(1.0 - b2);
    cudaStream_t stream = (cudaStream_t)stream_;
    const int slots = cfg->accumulation_steps;
    // Launch shape.  Two 512-thread CTAs per SM when their shared memory fits (227 KB/SM; static + 1 KB reserved per CTA =
    // 17 KB) with all slots' records resident, else one 1024-thread CTA per SM: 32 warps/SM either way.  The third shape --
    // two 512-thread CTAs with the records of two slots at a time (HALVES = 2) -- keeps two CTAs up to r_capacity 1024, but
    // measured 4 % SLOWER than the single 1024-thread CTA on the Panoptic shape (5 550 vs 5 784 frames/s, bit-identical
    // results; profiles/README.md) and 2-7 % slower where both fit, so it is only taken on request (resident_record_slots).
    auto fits2 = [&](int halves) { return SSB_OPT_MIN_CTAS * (opt_dyn_smem(slots, cfg->r_capacity, halves) + OPT_STATIC_SMEM) <= 227 * 1024; };
    if (cfg->resident_record_slots != 0 && cfg->resident_record_slots != slots && !(slots == 4 && cfg->resident_record_slots == 2)) return SSB_ERR_INVALID;
    const int halves = cfg->resident_record_slots ? slots / cfg->resident_record_slots : 1;
    const bool big = !fits2(halves);
    const size_t smem = opt_dyn_smem(slots, cfg->r_capacity, halves);
#define SSB_LAUNCH_OPT_K(S, NTH, HV)                                                                              \
    {                                                                                                             \
        if (cudaFuncSetAttribute(optimize_kernel<S, NTH, HV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) \
            return ssb_set_cuda_error(cudaGetLastError());                                                        \
        optimize_kernel<S, NTH, HV><<<n_frames, NTH, smem, stream>>>(p, tab);                                     \
    }
#define SSB_LAUNCH_OPT(S)                                                                                         \
    {                                                                                                             \
        if (big) SSB_LAUNCH_OPT_K(S, 1024, 1)                                                                     \
        else SSB_LAUNCH_OPT_K(S, OPT_THREADS, 1)                                                                  \
    }
    switch (slots) {
        case 1: SSB_LAUNCH_OPT(1) break;
        case 2: SSB_LAUNCH_OPT(2) break;
        case 4:
            if (halves == 2) SSB_LAUNCH_OPT_K(4, OPT_THREADS, 2)
            else SSB_LAUNCH_OPT(4)
            break;
        default: return SSB_ERR_UNSUPPORTED;
    }
    return ssb_set_cuda_error(cudaGetLastError());
}

int ssb_optimize_frames(const ssb_opt_config* cfg, int n_frames, const ssb_cameras* cams, const double* lr_xyz_host,
                        float* xyz, float* scaling_raw, float* rotation_raw, float* opacity_raw,
                        const int* roi_rect, const int64_t* roi_offset, const float* roi_data,
                        float* final_loss, void* workspace, void* stream_)
{
    return optimize_frames_impl(cfg, n_frames, cams, lr_xyz_host, xyz, scaling_raw, rotation_raw, opacity_raw, roi_rect, roi_offset,
                                roi_data, final_loss, workspace, -1, -1, nullptr, stream_);
}

int ssb_optimize_frames_debug(const ssb_opt_config* cfg, int n_frames, const ssb_cameras* cams, const double* lr_xyz_host,
                              float* xyz, float* scaling_raw, float* rotation_raw, float* opacity_raw,
                              const int* roi_rect, const int64_t* roi_offset, const float* roi_data,
                              float* final_loss, void* workspace, int dbg_frame, int dbg_step, int* dbg_out, void* stream_)
{
    if (!dbg_out || dbg_frame < 0 || dbg_frame >= n_frames || dbg_step < 0) return SSB_ERR_INVALID;
    return optimize_frames_impl(cfg, n_frames, cams, lr_xyz_host, xyz, scaling_raw, rotation_raw, opacity_raw, roi_rect, roi_offset,
                                roi_data, final_loss, workspace, dbg_frame, dbg_step, dbg_out, stream_);
}

}  // extern "C"
