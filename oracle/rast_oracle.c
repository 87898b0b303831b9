/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle, never on the product path.
 *
 * Plain-C restatement of the reference SkelSplat rasteriser (the modified 3DGS
 * rasteriser in submodules/diff-gaussian-rasterization-{h36m,panoptic,op}; the
 * three copies differ only in NUM_CHANNELS, which is a runtime argument here).
 * Citations are file:line under /root/reference/submodules/diff-gaussian-rasterization-h36m/ ("RAST/").
 *
 * Pinning status: the reference ships no golden vectors for the rasteriser
 * (SURVEY.md section 4).  This oracle is pinned against outputs of the compiled
 * reference itself (oracle/_ref, run on a B200 by tests/golden/make_golden.py ->
 * tests/golden/*.npz) -- see tests/test_oracle_golden.py.
 *
 * Floating point: compiled with -ffp-contract=off.  Wherever nvcc contracts the
 * reference's expression into an FMA (read off the PTX/SASS of the reference
 * built for sm_100a) fmaf() is written explicitly, so tile rectangles, depth
 * keys, sort order and ranges are reproduced bit for bit.  expf() is glibc's,
 * not CUDA's, so rendered values / gradients agree to ~1e-6 relative, not bitwise.
 * Backward accumulations (global atomicAdd in the reference, unordered) are done
 * in double here: the oracle is the "infinitely careful" sum.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define TILE 16 /* BLOCK_X == BLOCK_Y == 16, RAST/cuda_rasterizer/config.h:16-17 */

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* a*x + b*y + c*z + d exactly as nvcc emits it for transformPoint4x3/4x4
 * (RAST/cuda_rasterizer/auxiliary.h:70-89): t=b*y; t=fma(a,x,t); t=fma(c,z,t); t=d+t */
static inline float affine3(float a, float x, float b, float y, float c, float z, float d) {
    float t = b * y;
    t = fmaf(a, x, t);
    t = fmaf(c, z, t);
    return d + t;
}
/* dot of two 3-vectors as glm mat3*mat3 elements come out of nvcc:
 * fma(a2,b2, fma(a0,b0, a1*b1)) */
static inline float dot3(float a0, float b0, float a1, float b1, float a2, float b2) {
    float t = a1 * b1;
    t = fmaf(a0, b0, t);
    return fmaf(a2, b2, t);
}

/* ndc2Pix, RAST/cuda_rasterizer/auxiliary.h:40-43: evaluated in fp64 with one fma */
static inline float ndc2pix(float v, int S) {
    double d = fma((double)v + 1.0, (double)S, -1.0) * 0.5;
    return (float)d;
}

/* getRect, RAST/cuda_rasterizer/auxiliary.h:45-55 (division by 16 == multiply by 0.0625, exact) */
static void get_rect(float px, float py, int radius, int gx, int gy, uint32_t* r /*minx,miny,maxx,maxy*/) {
    float rf = (float)radius;
    r[0] = (uint32_t)imin(gx, imax(0, (int)((px - rf) * 0.0625f)));
    r[1] = (uint32_t)imin(gy, imax(0, (int)((py - rf) * 0.0625f)));
    r[2] = (uint32_t)imin(gx, imax(0, (int)((((px + rf) + 16.0f) + -1.0f) * 0.0625f)));
    r[3] = (uint32_t)imin(gy, imax(0, (int)((((py + rf) + 16.0f) + -1.0f) * 0.0625f)));
}

/* computeCov3D forward, RAST/cuda_rasterizer/forward.cu:114-150 (quaternion used un-normalised) */
static void cov3d_from_scale_rot(const float* scale, float mod, const float* q, float* cov) {
    float sx = mod * scale[0], sy = mod * scale[1], sz = mod * scale[2];
    float r = q[0], x = q[1], y = q[2], z = q[3];
    float yy = y * y, zz = z * z;
    float rz = r * z, xz = x * z, rx = r * x;
    /* ptxas fuses one product of each a*b +- c*d (read off the reference's SASS for sm_100a) */
    float R00 = 1.0f - ((yy + zz) + (yy + zz));
    float t01 = fmaf(x, y, -rz);  float R01 = t01 + t01;   /* xy - rz */
    float t02 = fmaf(r, y, xz);   float R02 = t02 + t02;   /* xz + ry */
    float t10 = fmaf(x, y, rz);   float R10 = t10 + t10;   /* xy + rz */
    float t11 = fmaf(x, x, zz);   float R11 = 1.0f - (t11 + t11);
    float t12 = fmaf(y, z, -rx);  float R12 = t12 + t12;   /* yz - rx */
    float t20 = fmaf(-r, y, xz);  float R20 = t20 + t20;   /* xz - ry */
    float t21 = fmaf(y, z, rx);   float R21 = t21 + t21;   /* yz + rx */
    float t22 = fmaf(x, x, yy);   float R22 = 1.0f - (t22 + t22);
    /* M = S*R in glm storage: column j = (sx*Rj0, sy*Rj1, sz*Rj2) */
    float A0 = sx * R00, A1 = sy * R01, A2 = sz * R02;
    float B0 = sx * R10, B1 = sy * R11, B2 = sz * R12;
    float C0 = sx * R20, C1 = sy * R21, C2 = sz * R22;
    cov[0] = dot3(A0, A0, A1, A1, A2, A2);
    cov[1] = dot3(B0, A0, B1, A1, B2, A2);
    cov[2] = dot3(C0, A0, C1, A1, C2, A2);
    cov[3] = dot3(B0, B0, B1, B1, B2, B2);
    cov[4] = dot3(C0, B0, C1, B1, C2, B2);
    cov[5] = dot3(C0, C0, C1, C1, C2, C2);
}

/*
 * preprocessCUDA forward, RAST/cuda_rasterizer/forward.cu:153-273.
 * Outputs are per Gaussian; entries of culled Gaussians (radii==0) keep their
 * previous contents except radii/tiles_touched (=0), as in the reference.
 * rects: [P][4] = minx,miny,maxx,maxy (not stored by the reference; exposed for tests).
 */
void oracle_preprocess(
    int P, const float* means3D, const float* scales, float scale_modifier, const float* rotations,
    const float* opacities, const float* cov3D_precomp, const float* view, const float* proj,
    int W, int H, float tan_fovx, float tan_fovy, int antialiasing,
    int* radii, float* means2D, float* depths, float* cov3Ds, float* conic_opacity,
    uint32_t* tiles_touched, uint32_t* rects)
{
    const float focal_y = H / (2.0f * tan_fovy); /* rasterizer_impl.cu:224-225 */
    const float focal_x = W / (2.0f * tan_fovx);
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    for (int i = 0; i < P; i++) {
        radii[i] = 0;
        tiles_touched[i] = 0;
        if (rects) memset(rects + 4 * i, 0, 16);
        float px = means3D[3 * i], py = means3D[3 * i + 1], pz = means3D[3 * i + 2];
        /* in_frustum, auxiliary.h:151-176: only the near test survives */
        float depth = affine3(view[2], px, view[6], py, view[10], pz, view[14]);
        if (depth <= 0.2f) continue;
        float hx = affine3(proj[0], px, proj[4], py, proj[8], pz, proj[12]);
        float hy = affine3(proj[1], px, proj[5], py, proj[9], pz, proj[13]);
        float hw = affine3(proj[3], px, proj[7], py, proj[11], pz, proj[15]);
        float p_w = 1.0f / (hw + 0.0000001f);
        float projx = hx * p_w, projy = hy * p_w;

        const float* cov3D;
        if (cov3D_precomp) {
            cov3D = cov3D_precomp + 6 * i;
        } else {
            cov3d_from_scale_rot(scales + 3 * i, scale_modifier, rotations + 4 * i, cov3Ds + 6 * i);
            cov3D = cov3Ds + 6 * i;
        }
        /* computeCov2D, forward.cu:74-109 */
        float tx = affine3(view[0], px, view[4], py, view[8], pz, view[12]);
        float ty = affine3(view[1], px, view[5], py, view[9], pz, view[13]);
        float tz = affine3(view[2], px, view[6], py, view[10], pz, view[14]);
        float limx = tan_fovx * 1.3f, limy = tan_fovy * 1.3f;
        float txtz = tx / tz, tytz = ty / tz;
        float cx = fminf(limx, fmaxf(-limx, txtz));
        float cy = fminf(limy, fmaxf(-limy, tytz));
        float tz2 = tz * tz;
        float J00 = focal_x / tz;
        float J02 = (focal_x * (cx * -tz)) / tz2; /* -(fx*t.x)/(tz*tz), t.x = clamp*tz */
        float J11 = focal_y / tz;
        float J12 = (focal_y * (cy * -tz)) / tz2;
        /* T = W*J, only two non-zero columns a,b */
        float a0 = fmaf(view[2], J02, view[0] * J00);
        float a1 = fmaf(view[6], J02, view[4] * J00);
        float a2 = fmaf(J02, view[10], view[8] * J00);
        float b0 = fmaf(view[2], J12, J11 * view[1]);
        float b1 = fmaf(view[6], J12, J11 * view[5]);
        float b2 = fmaf(J12, view[10], J11 * view[9]);
        float c0 = cov3D[0], c1 = cov3D[1], c2 = cov3D[2], c3 = cov3D[3], c4 = cov3D[4], c5 = cov3D[5];
        float va0 = dot3(a0, c0, a1, c1, a2, c2);
        float vb0 = dot3(b0, c0, b1, c1, b2, c2);
        float va1 = dot3(a0, c1, a1, c3, a2, c4);
        float vb1 = dot3(b0, c1, b1, c3, b2, c4);
        float va2 = dot3(a0, c2, a1, c4, a2, c5);
        float vb2 = dot3(b0, c2, b1, c4, b2, c5);
        float cov_x = dot3(a0, va0, a1, va1, a2, va2);
        float cov_y = dot3(a0, vb0, a1, vb1, a2, vb2);
        float cov_z = dot3(b0, vb0, b1, vb1, b2, vb2);

        float cyy = cov_y * cov_y;
        float det_cov = fmaf(cov_x, cov_z, -cyy);
        cov_x = cov_x + 0.3f;
        cov_z = cov_z + 0.3f;
        float det = fmaf(cov_x, cov_z, -cyy);
        float h_scaling = 1.0f;
        if (antialiasing) h_scaling = sqrtf(fmaxf(0.000025f, det_cov / det));
        if (det == 0.0f) continue;
        float det_inv = 1.0f / det;
        float conx = cov_z * det_inv, cony = det_inv * -cov_y, conz = cov_x * det_inv;
        float mid = (cov_x + cov_z) * 0.5f;
        float root = sqrtf(fmaxf(fmaf(mid, mid, -det), 0.1f));
        float lambda1 = mid + root, lambda2 = mid - root;
        float my_radius = ceilf(sqrtf(fmaxf(lambda1, lambda2)) * 3.0f);
        float pix_x = ndc2pix(projx, W), pix_y = ndc2pix(projy, H);
        int rad = (int)my_radius;
        uint32_t r[4];
        get_rect(pix_x, pix_y, rad, gx, gy, r);
        uint32_t tiles = (r[2] - r[0]) * (r[3] - r[1]);
        if (tiles == 0) continue;
        depths[i] = depth;
        radii[i] = rad;
        means2D[2 * i] = pix_x;
        means2D[2 * i + 1] = pix_y;
        conic_opacity[4 * i] = conx;
        conic_opacity[4 * i + 1] = cony;
        conic_opacity[4 * i + 2] = conz;
        conic_opacity[4 * i + 3] = h_scaling * opacities[i];
        tiles_touched[i] = tiles;
        if (rects) memcpy(rects + 4 * i, r, 16);
    }
}

/* getHigherMsb, RAST/cuda_rasterizer/rasterizer_impl.cu:35-50 */
uint32_t oracle_higher_msb(uint32_t n) {
    uint32_t msb = sizeof(n) * 4, step = msb;
    while (step > 1) {
        step /= 2;
        if (n >> msb) msb += step; else msb -= step;
    }
    if (n >> msb) msb++;
    return msb;
}

typedef struct { uint64_t key; uint32_t val; uint32_t pos; } kv_t;
static uint64_t g_sort_mask;
static int kv_cmp(const void* a, const void* b) {
    const kv_t* x = (const kv_t*)a; const kv_t* y = (const kv_t*)b;
    uint64_t kx = x->key & g_sort_mask, ky = y->key & g_sort_mask;
    if (kx != ky) return kx < ky ? -1 : 1;
    return x->pos < y->pos ? -1 : (x->pos > y->pos ? 1 : 0); /* stable (LSD radix sort) */
}

/*
 * Binning: inclusive scan (rasterizer_impl.cu:280), duplicateWithKeys (:70-111),
 * stable sort on the low 32+bit key bits (:303-311), memset + identifyTileRanges (:313-320,116-138).
 * Returns R. Arrays sized for the caller-computed R (= sum tiles_touched).
 */
int oracle_bin(int P, int W, int H, const int* radii, const float* means2D, const float* depths,
               const uint32_t* tiles_touched, uint32_t* point_offsets,
               uint64_t* keys_unsorted, uint32_t* vals_unsorted,
               uint64_t* keys_sorted, uint32_t* vals_sorted, uint32_t* ranges /*[tiles][2]*/)
{
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    uint32_t acc = 0;
    for (int i = 0; i < P; i++) { acc += tiles_touched[i]; point_offsets[i] = acc; }
    int R = (int)acc;
    for (int i = 0; i < P; i++) {
        if (radii[i] > 0) {
            uint32_t off = (i == 0) ? 0 : point_offsets[i - 1];
            uint32_t r[4];
            get_rect(means2D[2 * i], means2D[2 * i + 1], radii[i], gx, gy, r);
            for (uint32_t y = r[1]; y < r[3]; y++)
                for (uint32_t x = r[0]; x < r[2]; x++) {
                    uint64_t key = (uint64_t)(y * (uint32_t)gx + x);
                    key <<= 32;
                    key |= f2u(depths[i]);
                    keys_unsorted[off] = key;
                    vals_unsorted[off] = (uint32_t)i;
                    off++;
                }
        }
    }
    uint32_t bit = oracle_higher_msb((uint32_t)(gx * gy));
    kv_t* kv = (kv_t*)malloc(sizeof(kv_t) * (size_t)(R > 0 ? R : 1));
    for (int i = 0; i < R; i++) { kv[i].key = keys_unsorted[i]; kv[i].val = vals_unsorted[i]; kv[i].pos = (uint32_t)i; }
    g_sort_mask = (32 + bit >= 64) ? ~0ull : ((1ull << (32 + bit)) - 1ull);
    qsort(kv, (size_t)R, sizeof(kv_t), kv_cmp);
    for (int i = 0; i < R; i++) { keys_sorted[i] = kv[i].key; vals_sorted[i] = kv[i].val; }
    free(kv);
    memset(ranges, 0, sizeof(uint32_t) * 2 * (size_t)(gx * gy));
    for (int i = 0; i < R; i++) {
        uint32_t cur = (uint32_t)(keys_sorted[i] >> 32);
        if (i == 0) ranges[2 * cur] = 0;
        else {
            uint32_t prev = (uint32_t)(keys_sorted[i - 1] >> 32);
            if (cur != prev) { ranges[2 * prev + 1] = (uint32_t)i; ranges[2 * cur] = (uint32_t)i; }
        }
        if (i == R - 1) ranges[2 * cur + 1] = (uint32_t)R;
    }
    return R;
}

/* alpha of one (pixel, Gaussian) pair: power with the reference's operation order
 * (forward.cu:352-362). Returns 0 if the pair is skipped. */
static inline int pair_alpha(float gx_, float gy_, float pxf, float pyf, const float* con_o,
                             float* dx_o, float* dy_o, float* G_o, float* alpha_o) {
    float dx = gx_ - pxf, dy = gy_ - pyf;
    float t = dy * (dy * con_o[2]);
    t = fmaf(dx, dx * con_o[0], t);
    float power = fmaf(t, -0.5f, -(dy * (dx * con_o[1])));
    if (power > 0.0f) return 0;
    float G = expf(power);
    float alpha = fminf(0.99f, con_o[3] * G);
    if (alpha < 1.0f / 255.0f) return 0;
    *dx_o = dx; *dy_o = dy; *G_o = G; *alpha_o = alpha;
    return 1;
}

/*
 * renderCUDA forward, RAST/cuda_rasterizer/forward.cu:278-401.  features = [P][C]
 * (the reference passes `shs` here, rasterizer_impl.cu:324-331).  Background is not
 * blended (forward.cu:396).  out_color must be pre-zeroed by the caller only for
 * symmetry: every pixel is written.
 */
void oracle_render_forward(int C, int W, int H, const uint32_t* ranges, const uint32_t* point_list,
                           const float* means2D, const float* features, const float* conic_opacity,
                           const float* depths, float* final_T, uint32_t* n_contrib,
                           float* out_color, float* invdepth)
{
    const int gx = (W + TILE - 1) / TILE;
    const size_t HW = (size_t)H * W;
    float* acc = (float*)malloc(sizeof(float) * (size_t)C);
    for (int y = 0; y < H; y++) {
        for (int x = 0; x < W; x++) {
            int tile = (y / TILE) * gx + (x / TILE);
            uint32_t s = ranges[2 * tile], e = ranges[2 * tile + 1];
            float T = 1.0f, inv = 0.0f;
            uint32_t contributor = 0, last = 0;
            for (int c = 0; c < C; c++) acc[c] = 0.0f;
            for (uint32_t k = s; k < e; k++) {
                contributor++;
                uint32_t id = point_list[k];
                float dx, dy, G, alpha;
                if (!pair_alpha(means2D[2 * id], means2D[2 * id + 1], (float)x, (float)y,
                                conic_opacity + 4 * id, &dx, &dy, &G, &alpha)) continue;
                float test_T = T * (1.0f - alpha);
                if (test_T < 0.0001f) break; /* done=true: no later Gaussian is touched */
                for (int c = 0; c < C; c++) acc[c] = fmaf(T, alpha * features[(size_t)id * C + c], acc[c]);
                inv = fmaf(T, alpha * (1.0f / depths[id]), inv);
                T = test_T;
                last = contributor;
            }
            size_t pix = (size_t)y * W + x;
            final_T[pix] = T;
            n_contrib[pix] = last;
            for (int c = 0; c < C; c++) out_color[(size_t)c * HW + pix] = acc[c];
            if (invdepth) invdepth[pix] = inv;
        }
    }
    free(acc);
}

/*
 * renderCUDA backward, RAST/cuda_rasterizer/backward.cu:452-638.  bg is defined as 0
 * (the reference reads past a 3-float bg tensor; contents are zeros).
 * Accumulators are double[P][...]; outputs are written as float.
 * dL_dmean2D [P][3] (z unused), dL_dconic [P][4] (x,y,-,w), dL_dopacity [P], dL_dcolors [P][C], dL_dinvdepths [P].
 */
void oracle_render_backward(int P, int C, int W, int H, const uint32_t* ranges, const uint32_t* point_list,
                            const float* means2D, const float* conic_opacity, const float* colors,
                            const float* depths, const float* final_Ts, const uint32_t* n_contrib,
                            const float* dL_dpixels, const float* dL_dinvdepth_pix,
                            float* dL_dmean2D, float* dL_dconic, float* dL_dopacity,
                            float* dL_dcolors, float* dL_dinvdepths)
{
    const int gx = (W + TILE - 1) / TILE;
    const size_t HW = (size_t)H * W;
    const int S = 7 + C;
    double* g = (double*)calloc((size_t)P * S, sizeof(double));
    float* accum_rec = (float*)malloc(sizeof(float) * C);
    float* last_color = (float*)malloc(sizeof(float) * C);
    float* dpix = (float*)malloc(sizeof(float) * C);
    const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
    for (int y = 0; y < H; y++) {
        for (int x = 0; x < W; x++) {
            int tile = (y / TILE) * gx + (x / TILE);
            uint32_t s = ranges[2 * tile], e = ranges[2 * tile + 1];
            if (e == s) continue;
            size_t pix = (size_t)y * W + x;
            const float T_final = final_Ts[pix];
            float T = T_final;
            uint32_t contributor = e - s;
            const uint32_t last_contributor = n_contrib[pix];
            for (int c = 0; c < C; c++) { accum_rec[c] = 0.f; last_color[c] = 0.f; dpix[c] = dL_dpixels[(size_t)c * HW + pix]; }
            float dinv_pix = dL_dinvdepth_pix ? dL_dinvdepth_pix[pix] : 0.f;
            float last_alpha = 0.f, last_invdepth = 0.f, accum_invdepth_rec = 0.f;
            for (uint32_t k = e; k-- > s;) {
                contributor--;
                if (contributor >= last_contributor) continue;
                uint32_t id = point_list[k];
                const float* con_o = conic_opacity + 4 * id;
                float dx, dy, G, alpha;
                if (!pair_alpha(means2D[2 * id], means2D[2 * id + 1], (float)x, (float)y, con_o, &dx, &dy, &G, &alpha)) continue;
                T = T / (1.f - alpha);
                const float dchannel_dcolor = alpha * T;
                float dL_dalpha = 0.f;
                double* gi = g + (size_t)id * S;
                for (int c = 0; c < C; c++) {
                    const float col = colors[(size_t)id * C + c];
                    accum_rec[c] = last_alpha * last_color[c] + (1.f - last_alpha) * accum_rec[c];
                    last_color[c] = col;
                    dL_dalpha += (col - accum_rec[c]) * dpix[c];
                    gi[7 + c] += (double)(dchannel_dcolor * dpix[c]);
                }
                if (dL_dinvdepth_pix) {
                    const float invd = 1.f / depths[id];
                    accum_invdepth_rec = last_alpha * last_invdepth + (1.f - last_alpha) * accum_invdepth_rec;
                    last_invdepth = invd;
                    dL_dalpha += (invd - accum_invdepth_rec) * dinv_pix;
                    gi[6] += (double)(dchannel_dcolor * dinv_pix);
                }
                dL_dalpha *= T;
                last_alpha = alpha;
                /* bg term: bg == 0 */
                const float dL_dG = con_o[3] * dL_dalpha;
                const float gdx = G * dx, gdy = G * dy;
                const float dG_ddelx = -gdx * con_o[0] - gdy * con_o[1];
                const float dG_ddely = -gdy * con_o[2] - gdx * con_o[1];
                gi[0] += (double)(dL_dG * dG_ddelx * ddelx_dx);
                gi[1] += (double)(dL_dG * dG_ddely * ddely_dy);
                gi[2] += (double)(-0.5f * gdx * dx * dL_dG);
                gi[3] += (double)(-0.5f * gdx * dy * dL_dG);
                gi[4] += (double)(-0.5f * gdy * dy * dL_dG);
                gi[5] += (double)(G * dL_dalpha);
            }
        }
    }
    for (int i = 0; i < P; i++) {
        const double* gi = g + (size_t)i * S;
        dL_dmean2D[3 * i] = (float)gi[0]; dL_dmean2D[3 * i + 1] = (float)gi[1]; dL_dmean2D[3 * i + 2] = 0.f;
        dL_dconic[4 * i] = (float)gi[2]; dL_dconic[4 * i + 1] = (float)gi[3]; dL_dconic[4 * i + 2] = 0.f; dL_dconic[4 * i + 3] = (float)gi[4];
        dL_dopacity[i] = (float)gi[5];
        if (dL_dinvdepths) dL_dinvdepths[i] = (float)gi[6];
        for (int c = 0; c < C; c++) dL_dcolors[(size_t)i * C + c] = (float)gi[7 + c];
    }
    free(g); free(accum_rec); free(last_color); free(dpix);
}

/*
 * BACKWARD::preprocess = computeCov2DCUDA (backward.cu:147-326) then preprocessCUDA
 * (backward.cu:398-449) with computeCov3D backward (backward.cu:330-393).
 * antialiasing == false path only (all shipped configs).
 * dL_dinvdepth may be NULL. dL_dmean3D is ASSIGNED by the first stage then
 * accumulated by the second, as in the reference.  Gaussians with radii<=0 keep zeros.
 */
void oracle_preprocess_backward(
    int P, const float* means3D, const int* radii, const float* cov3Ds, const float* scales,
    const float* rotations, float scale_modifier, const float* view, const float* proj,
    int W, int H, float tan_fovx, float tan_fovy,
    const float* dL_dmean2D, const float* dL_dconics, const float* dL_dinvdepth,
    float* dL_dmean3D, float* dL_dcov3D, float* dL_dscale, float* dL_drot)
{
    const float h_y = H / (2.0f * tan_fovy);
    const float h_x = W / (2.0f * tan_fovx);
    for (int idx = 0; idx < P; idx++) {
        if (!(radii[idx] > 0)) continue;
        const float* cov3D = cov3Ds + 6 * idx;
        float mx = means3D[3 * idx], my = means3D[3 * idx + 1], mz = means3D[3 * idx + 2];
        float dconx = dL_dconics[4 * idx], dcony = dL_dconics[4 * idx + 1], dconz = dL_dconics[4 * idx + 3];
        float tx = view[0] * mx + view[4] * my + view[8] * mz + view[12];
        float ty = view[1] * mx + view[5] * my + view[9] * mz + view[13];
        float tz = view[2] * mx + view[6] * my + view[10] * mz + view[14];
        const float limx = 1.3f * tan_fovx, limy = 1.3f * tan_fovy;
        const float txtz = tx / tz, tytz = ty / tz;
        tx = fminf(limx, fmaxf(-limx, txtz)) * tz;
        ty = fminf(limy, fmaxf(-limy, tytz)) * tz;
        const float x_grad_mul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
        const float y_grad_mul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
        /* glm storage: M[c][r]. J cols: (hx/tz,0,-(hx tx)/tz^2), (0,hy/tz,-(hy ty)/tz^2), 0 */
        float J[3][3] = {{h_x / tz, 0.f, -(h_x * tx) / (tz * tz)}, {0.f, h_y / tz, -(h_y * ty) / (tz * tz)}, {0.f, 0.f, 0.f}};
        float Wm[3][3] = {{view[0], view[4], view[8]}, {view[1], view[5], view[9]}, {view[2], view[6], view[10]}};
        float Vrk[3][3] = {{cov3D[0], cov3D[1], cov3D[2]}, {cov3D[1], cov3D[3], cov3D[4]}, {cov3D[2], cov3D[4], cov3D[5]}};
        float T[3][3]; /* T = W*J (glm): T[c][r] = sum_k W[k][r]*J[c][k] */
        for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) {
            float s = 0.f; for (int k = 0; k < 3; k++) s += Wm[k][r] * J[c][k]; T[c][r] = s; }
        /* cov2D = T^T * Vrk^T * T: cov2D[c][r] = sum_{i,j} T[r][i] Vrk[i][j] T[c][j]   (glm index algebra) */
        float c_xx = 0.f, c_xy = 0.f, c_yy = 0.f;
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
            c_xx += T[0][i] * Vrk[i][j] * T[0][j];
            c_xy += T[0][i] * Vrk[i][j] * T[1][j];
            c_yy += T[1][i] * Vrk[i][j] * T[1][j];
        }
        c_xx += 0.3f; c_yy += 0.3f;
        float dL_dc_xx = 0.f, dL_dc_xy = 0.f, dL_dc_yy = 0.f;
        float denom = c_xx * c_yy - c_xy * c_xy;
        float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
        float* dcov = dL_dcov3D + 6 * idx;
        if (denom2inv != 0) {
            dL_dc_xx += denom2inv * (-c_yy * c_yy * dconx + 2 * c_xy * c_yy * dcony + (denom - c_xx * c_yy) * dconz);
            dL_dc_yy += denom2inv * (-c_xx * c_xx * dconz + 2 * c_xx * c_xy * dcony + (denom - c_xx * c_yy) * dconx);
            dL_dc_xy += denom2inv * 2 * (c_xy * c_yy * dconx - (denom + 2 * c_xy * c_xy) * dcony + c_xx * c_xy * dconz);
            dcov[0] = (T[0][0] * T[0][0] * dL_dc_xx + T[0][0] * T[1][0] * dL_dc_xy + T[1][0] * T[1][0] * dL_dc_yy);
            dcov[3] = (T[0][1] * T[0][1] * dL_dc_xx + T[0][1] * T[1][1] * dL_dc_xy + T[1][1] * T[1][1] * dL_dc_yy);
            dcov[5] = (T[0][2] * T[0][2] * dL_dc_xx + T[0][2] * T[1][2] * dL_dc_xy + T[1][2] * T[1][2] * dL_dc_yy);
            dcov[1] = 2 * T[0][0] * T[0][1] * dL_dc_xx + (T[0][0] * T[1][1] + T[0][1] * T[1][0]) * dL_dc_xy + 2 * T[1][0] * T[1][1] * dL_dc_yy;
            dcov[2] = 2 * T[0][0] * T[0][2] * dL_dc_xx + (T[0][0] * T[1][2] + T[0][2] * T[1][0]) * dL_dc_xy + 2 * T[1][0] * T[1][2] * dL_dc_yy;
            dcov[4] = 2 * T[0][2] * T[0][1] * dL_dc_xx + (T[0][1] * T[1][2] + T[0][2] * T[1][1]) * dL_dc_xy + 2 * T[1][1] * T[1][2] * dL_dc_yy;
        } else {
            for (int i = 0; i < 6; i++) dcov[i] = 0;
        }
        float dL_dT00 = 2 * (T[0][0] * Vrk[0][0] + T[0][1] * Vrk[0][1] + T[0][2] * Vrk[0][2]) * dL_dc_xx + (T[1][0] * Vrk[0][0] + T[1][1] * Vrk[0][1] + T[1][2] * Vrk[0][2]) * dL_dc_xy;
        float dL_dT01 = 2 * (T[0][0] * Vrk[1][0] + T[0][1] * Vrk[1][1] + T[0][2] * Vrk[1][2]) * dL_dc_xx + (T[1][0] * Vrk[1][0] + T[1][1] * Vrk[1][1] + T[1][2] * Vrk[1][2]) * dL_dc_xy;
        float dL_dT02 = 2 * (T[0][0] * Vrk[2][0] + T[0][1] * Vrk[2][1] + T[0][2] * Vrk[2][2]) * dL_dc_xx + (T[1][0] * Vrk[2][0] + T[1][1] * Vrk[2][1] + T[1][2] * Vrk[2][2]) * dL_dc_xy;
        float dL_dT10 = 2 * (T[1][0] * Vrk[0][0] + T[1][1] * Vrk[0][1] + T[1][2] * Vrk[0][2]) * dL_dc_yy + (T[0][0] * Vrk[0][0] + T[0][1] * Vrk[0][1] + T[0][2] * Vrk[0][2]) * dL_dc_xy;
        float dL_dT11 = 2 * (T[1][0] * Vrk[1][0] + T[1][1] * Vrk[1][1] + T[1][2] * Vrk[1][2]) * dL_dc_yy + (T[0][0] * Vrk[1][0] + T[0][1] * Vrk[1][1] + T[0][2] * Vrk[1][2]) * dL_dc_xy;
        float dL_dT12 = 2 * (T[1][0] * Vrk[2][0] + T[1][1] * Vrk[2][1] + T[1][2] * Vrk[2][2]) * dL_dc_yy + (T[0][0] * Vrk[2][0] + T[0][1] * Vrk[2][1] + T[0][2] * Vrk[2][2]) * dL_dc_xy;
        float dL_dJ00 = Wm[0][0] * dL_dT00 + Wm[0][1] * dL_dT01 + Wm[0][2] * dL_dT02;
        float dL_dJ02 = Wm[2][0] * dL_dT00 + Wm[2][1] * dL_dT01 + Wm[2][2] * dL_dT02;
        float dL_dJ11 = Wm[1][0] * dL_dT10 + Wm[1][1] * dL_dT11 + Wm[1][2] * dL_dT12;
        float dL_dJ12 = Wm[2][0] * dL_dT10 + Wm[2][1] * dL_dT11 + Wm[2][2] * dL_dT12;
        float itz = 1.f / tz, itz2 = itz * itz, itz3 = itz2 * itz;
        float dL_dtx = x_grad_mul * -h_x * itz2 * dL_dJ02;
        float dL_dty = y_grad_mul * -h_y * itz2 * dL_dJ12;
        float dL_dtz = -h_x * itz2 * dL_dJ00 - h_y * itz2 * dL_dJ11 + (2 * h_x * tx) * itz3 * dL_dJ02 + (2 * h_y * ty) * itz3 * dL_dJ12;
        if (dL_dinvdepth) dL_dtz -= dL_dinvdepth[idx] / (tz * tz);
        /* transformVec4x3Transpose, auxiliary.h:101-109 */
        float gmx = view[0] * dL_dtx + view[1] * dL_dty + view[2] * dL_dtz;
        float gmy = view[4] * dL_dtx + view[5] * dL_dty + view[6] * dL_dtz;
        float gmz = view[8] * dL_dtx + view[9] * dL_dty + view[10] * dL_dtz;

        /* second kernel: projection part, backward.cu:424-440 */
        float m_w = 1.0f / ((proj[3] * mx + proj[7] * my + proj[11] * mz + proj[15]) + 0.0000001f);
        float mul1 = (proj[0] * mx + proj[4] * my + proj[8] * mz + proj[12]) * m_w * m_w;
        float mul2 = (proj[1] * mx + proj[5] * my + proj[9] * mz + proj[13]) * m_w * m_w;
        float d2x = dL_dmean2D[3 * idx], d2y = dL_dmean2D[3 * idx + 1];
        gmx += (proj[0] * m_w - proj[3] * mul1) * d2x + (proj[1] * m_w - proj[3] * mul2) * d2y;
        gmy += (proj[4] * m_w - proj[7] * mul1) * d2x + (proj[5] * m_w - proj[7] * mul2) * d2y;
        gmz += (proj[8] * m_w - proj[11] * mul1) * d2x + (proj[9] * m_w - proj[11] * mul2) * d2y;
        dL_dmean3D[3 * idx] = gmx; dL_dmean3D[3 * idx + 1] = gmy; dL_dmean3D[3 * idx + 2] = gmz;

        /* computeCov3D backward, backward.cu:330-393 */
        if (scales) {
            const float* q = rotations + 4 * idx;
            float r = q[0], x = q[1], y = q[2], z = q[3];
            /* glm R[c][r] */
            float R[3][3] = {
                {1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y)},
                {2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x)},
                {2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)}};
            float s[3] = {scale_modifier * scales[3 * idx], scale_modifier * scales[3 * idx + 1], scale_modifier * scales[3 * idx + 2]};
            float M[3][3]; /* M = S*R: M[c][r] = s[r]*R[c][r] */
            for (int c = 0; c < 3; c++) for (int rr = 0; rr < 3; rr++) M[c][rr] = s[rr] * R[c][rr];
            float dS[3][3] = {{dcov[0], 0.5f * dcov[1], 0.5f * dcov[2]}, {0.5f * dcov[1], dcov[3], 0.5f * dcov[4]}, {0.5f * dcov[2], 0.5f * dcov[4], dcov[5]}};
            float dM[3][3]; /* dL_dM = 2*M*dSigma: dM[c][r] = 2*sum_k M[k][r]*dS[c][k] */
            for (int c = 0; c < 3; c++) for (int rr = 0; rr < 3; rr++) {
                float a = 0.f; for (int k = 0; k < 3; k++) a += M[k][rr] * dS[c][k]; dM[c][rr] = 2.0f * a; }
            /* Rt[c][r] = R[r][c]; dMt[c][r] = dM[r][c] */
            float dMt[3][3];
            for (int c = 0; c < 3; c++) for (int rr = 0; rr < 3; rr++) dMt[c][rr] = dM[rr][c];
            for (int c = 0; c < 3; c++) {
                float a = 0.f; for (int k = 0; k < 3; k++) a += R[k][c] * dMt[c][k];
                dL_dscale[3 * idx + c] = a;
            }
            for (int k = 0; k < 3; k++) { dMt[0][k] *= s[0]; dMt[1][k] *= s[1]; dMt[2][k] *= s[2]; }
            dL_drot[4 * idx + 0] = 2 * z * (dMt[0][1] - dMt[1][0]) + 2 * y * (dMt[2][0] - dMt[0][2]) + 2 * x * (dMt[1][2] - dMt[2][1]);
            dL_drot[4 * idx + 1] = 2 * y * (dMt[1][0] + dMt[0][1]) + 2 * z * (dMt[2][0] + dMt[0][2]) + 2 * r * (dMt[1][2] - dMt[2][1]) - 4 * x * (dMt[2][2] + dMt[1][1]);
            dL_drot[4 * idx + 2] = 2 * x * (dMt[1][0] + dMt[0][1]) + 2 * r * (dMt[2][0] - dMt[0][2]) + 2 * z * (dMt[1][2] + dMt[2][1]) - 4 * y * (dMt[2][2] + dMt[0][0]);
            dL_drot[4 * idx + 3] = 2 * r * (dMt[0][1] - dMt[1][0]) + 2 * x * (dMt[2][0] + dMt[0][2]) + 2 * y * (dMt[1][2] + dMt[2][1]) - 4 * z * (dMt[1][1] + dMt[0][0]);
        }
    }
}

/* checkFrustum / markVisible, rasterizer_impl.cu:54-66 */
void oracle_mark_visible(int P, const float* means3D, const float* view, uint8_t* present) {
    for (int i = 0; i < P; i++) {
        float depth = affine3(view[2], means3D[3 * i], view[6], means3D[3 * i + 1], view[10], means3D[3 * i + 2], view[14]);
        present[i] = depth > 0.2f;
    }
}
