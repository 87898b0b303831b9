// Dense-contract batched rasteriser (the drop-in behind GaussianRasterizer), sm_100a.
//
// Contract (reference: RAST/rasterize_points.cu:35-223): forward returns a freshly written
// [C,H,W] image + [1,H,W] inverse depth + radii; backward consumes dense dL/dimage.
// Design (B200-first):
//   * bin_kernel      one CTA per view: EWA projection of the view's few Gaussians, (tile|depth)
//                     keys, in-shared-memory bitonic sort, tile ranges + compact active-tile list.
//                     Replaces preprocessCUDA + cub scan + D2H sync + duplicateWithKeys + cub radix
//                     sort + memset + identifyTileRanges (6 launches, 1 host sync) by 1 launch.
//   * fill_zero + render_active: the dense image is streamed out as zeros with 128-bit stores (pure
//                     HBM-write stream) and only the ~100 ACTIVE tiles are composited and overwritten.
//                     No final_T / n_contrib side buffers: backward recomputes them per active tile.
//   * render_bwd      CTAs loop over ACTIVE tiles only; reads dL/dimage on those tiles only;
//                     per-(tile,Gaussian) partial sums by warp shuffles -> scratch, no atomics.
//   * gauss_bwd       one CTA per view: fixed-order sum of the partials + EWA / projection / cov3D chain.
#include "common.cuh"
#include "api_internal.h"

namespace ssb {

// ------------------------------------------------------------------------------------------ layout
__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

__host__ __device__ inline StateLayout state_layout(int P, int W, int H, int rcap) {
    StateLayout L;
    const size_t tiles = (size_t)((W + TILE - 1) / TILE) * ((H + TILE - 1) / TILE);
    size_t o = 0;
    L.off[SSB_F_HEADER] = o;         o = align_up(o + 8 * 4, 128);
    L.off[SSB_F_DEPTHS] = o;         o = align_up(o + (size_t)P * 4, 128);
    L.off[SSB_F_MEANS2D] = o;        o = align_up(o + (size_t)P * 8, 128);
    L.off[SSB_F_CONIC_OPACITY] = o;  o = align_up(o + (size_t)P * 16, 128);
    L.off[SSB_F_COV3D] = o;          o = align_up(o + (size_t)P * 24, 128);
    L.off[SSB_F_TILES_TOUCHED] = o;  o = align_up(o + (size_t)P * 4, 128);
    L.off[SSB_F_POINT_OFFSETS] = o;  o = align_up(o + (size_t)P * 4, 128);
    L.off[SSB_F_RECTS] = o;          o = align_up(o + (size_t)P * 16, 128);
    L.off[SSB_F_KEYS_UNSORTED] = o;  o = align_up(o + (size_t)rcap * 8, 128);
    L.off[SSB_F_VALS_UNSORTED] = o;  o = align_up(o + (size_t)rcap * 4, 128);
    L.off[SSB_F_KEYS_SORTED] = o;    o = align_up(o + (size_t)rcap * 8, 128);
    L.off[SSB_F_POINT_LIST] = o;     o = align_up(o + (size_t)rcap * 4, 128);
    L.off[SSB_F_INV_POS] = o;        o = align_up(o + (size_t)rcap * 4, 128);
    L.off[SSB_F_TILE_IDS] = o;       o = align_up(o + (size_t)rcap * 4, 128);
    L.off[SSB_F_TILE_RANGES] = o;    o = align_up(o + (size_t)rcap * 8, 128);
    L.off[SSB_F_RANGES] = o;         o = align_up(o + tiles * 8, 128);
    L.total = align_up(o, 512);
    return L;
}

struct ViewInfo {
    int W, H, gx, gy;
    float tan_fovx, tan_fovy, focal_x, focal_y;
};

__device__ __forceinline__ ViewInfo view_info(const ssb_cameras& cams, int cam) {
    ViewInfo v;
    v.W = cams.dims ? cams.dims[2 * cam] : cams.W0;
    v.H = cams.dims ? cams.dims[2 * cam + 1] : cams.H0;
    v.tan_fovx = cams.tanfov ? cams.tanfov[2 * cam] : cams.tanfovx0;
    v.tan_fovy = cams.tanfov ? cams.tanfov[2 * cam + 1] : cams.tanfovy0;
    v.gx = (v.W + TILE - 1) / TILE;
    v.gy = (v.H + TILE - 1) / TILE;
    // rasterizer_impl.cu:224-225
    v.focal_y = __fdiv_rn((float)v.H, __fmul_rn(2.0f, v.tan_fovy));
    v.focal_x = __fdiv_rn((float)v.W, __fmul_rn(2.0f, v.tan_fovx));
    return v;
}

template <typename T>
__device__ __forceinline__ T* field(char* base, const StateLayout& L, int f) { return reinterpret_cast<T*>(base + L.off[f]); }
template <typename T>
__device__ __forceinline__ const T* cfield(const char* base, const StateLayout& L, int f) { return reinterpret_cast<const T*>(base + L.off[f]); }

// ------------------------------------------------------------------------------------------ bin
constexpr int BIN_THREADS = 256;

// dynamic smem: keys u64[n2] | vals u32[n2] | scan u32[P] | misc
__global__ void __launch_bounds__(BIN_THREADS)
bin_kernel(ssb_gaussians g, ssb_cameras cams, int rcap, int n2, StateLayout L, int Wmax, int Hmax,
           char* __restrict__ state, int* __restrict__ radii_out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* s_keys = reinterpret_cast<uint64_t*>(smem_raw);
    uint32_t* s_vals = reinterpret_cast<uint32_t*>(s_keys + n2);
    uint32_t* s_scan = s_vals + n2;
    __shared__ float s_view[16], s_proj[16];
    __shared__ int s_R, s_nact;

    const int b = blockIdx.x;
    const int frame = b / cams.n_views, cam = b % cams.n_views;
    const int P = g.P;
    const int tid = threadIdx.x;
    char* st = state + (size_t)b * L.total;
    if (tid < 16) s_view[tid] = cams.viewmatrix[16 * cam + tid];
    else if (tid < 32) s_proj[tid - 16] = cams.projmatrix[16 * cam + tid - 16];
    const ViewInfo vi = view_info(cams, cam);
    __syncthreads();

    float* depths = field<float>(st, L, SSB_F_DEPTHS);
    float2* means2D = field<float2>(st, L, SSB_F_MEANS2D);
    float4* conic_opacity = field<float4>(st, L, SSB_F_CONIC_OPACITY);
    float* cov3Ds = field<float>(st, L, SSB_F_COV3D);
    uint32_t* tiles_touched = field<uint32_t>(st, L, SSB_F_TILES_TOUCHED);
    uint32_t* offsets = field<uint32_t>(st, L, SSB_F_POINT_OFFSETS);
    uint4* rects = field<uint4>(st, L, SSB_F_RECTS);

    // ---- per-Gaussian projection (preprocessCUDA)
    for (int i = tid; i < P; i += blockDim.x) {
        const size_t gi = (size_t)frame * P + i;
        const float mx = g.means3D[3 * gi], my = g.means3D[3 * gi + 1], mz = g.means3D[3 * gi + 2];
        float cov[6];
        if (g.cov3D_precomp) {
#pragma unroll
            for (int k = 0; k < 6; k++) cov[k] = g.cov3D_precomp[6 * gi + k];
        } else {
            cov3d_from_scale_rot(g.scales[3 * gi], g.scales[3 * gi + 1], g.scales[3 * gi + 2], g.scale_modifier,
                                 g.rotations[4 * gi], g.rotations[4 * gi + 1], g.rotations[4 * gi + 2], g.rotations[4 * gi + 3], cov);
        }
        const Splat s = project_gaussian(mx, my, mz, cov, g.opacities[gi], s_view, s_proj, vi.W, vi.H,
                                         vi.tan_fovx, vi.tan_fovy, vi.focal_x, vi.focal_y, cams.antialiasing != 0);
        depths[i] = s.depth;
        means2D[i] = make_float2(s.px, s.py);
        conic_opacity[i] = make_float4(s.conx, s.cony, s.conz, s.opac);
#pragma unroll
        for (int k = 0; k < 6; k++) cov3Ds[6 * i + k] = cov[k];
        tiles_touched[i] = s.tiles;
        rects[i] = make_uint4(s.rect.x0, s.rect.y0, s.rect.x1, s.rect.y1);
        radii_out[(size_t)b * P + i] = s.radius;
        s_scan[i] = s.tiles;
    }
    __syncthreads();
    // ---- inclusive scan of tiles_touched (cub::DeviceScan::InclusiveSum in the reference)
    for (int d = 1; d < P; d <<= 1) {
        uint32_t add[4];
        int cnt = 0;
        for (int i = tid; i < P; i += blockDim.x) add[cnt++] = (i >= d) ? s_scan[i - d] : 0u;
        __syncthreads();
        cnt = 0;
        for (int i = tid; i < P; i += blockDim.x) s_scan[i] += add[cnt++];
        __syncthreads();
    }
    for (int i = tid; i < P; i += blockDim.x) offsets[i] = s_scan[i];
    const uint32_t R_full = P > 0 ? s_scan[P - 1] : 0u;
    const uint32_t R = min(R_full, (uint32_t)rcap);

    // ---- duplicateWithKeys: emission order is Gaussian-major, then row-major tiles
    uint64_t* keys_unsorted = field<uint64_t>(st, L, SSB_F_KEYS_UNSORTED);
    uint32_t* vals_unsorted = field<uint32_t>(st, L, SSB_F_VALS_UNSORTED);
    // sort size: smallest power of two >= R (uniform across the CTA), at least one warp's worth
    int nsort = 32;
    while (nsort < (int)R) nsort <<= 1;
    if (nsort > n2) nsort = n2;
    for (int i = tid; i < nsort; i += blockDim.x) { s_keys[i] = ~0ull; s_vals[i] = 0xFFFFFFFFu; }
    __syncthreads();
    for (int i = tid; i < P; i += blockDim.x) {
        const uint32_t tiles = tiles_touched[i];
        if (tiles == 0) continue;
        uint32_t off = (i == 0) ? 0u : s_scan[i - 1];
        const uint4 r = rects[i];
        const uint32_t dbits = __float_as_uint(depths[i]);
        for (uint32_t y = r.y; y < r.w; y++)
            for (uint32_t x = r.x; x < r.z; x++) {
                if (off < R) {
                    const uint64_t key = ((uint64_t)(y * (uint32_t)vi.gx + x) << 32) | dbits;
                    s_keys[off] = key;
                    s_vals[off] = (off << 10) | (uint32_t)i;   // emission index breaks ties => stable
                    keys_unsorted[off] = key;
                    vals_unsorted[off] = (uint32_t)i;
                }
                off++;
            }
    }
    __syncthreads();
    // ---- sort (cub::DeviceRadixSort::SortPairs over bits [0, 32+bit) in the reference)
    bitonic_sort_cta(s_keys, s_vals, nsort);

    uint64_t* keys_sorted = field<uint64_t>(st, L, SSB_F_KEYS_SORTED);
    uint32_t* point_list = field<uint32_t>(st, L, SSB_F_POINT_LIST);
    uint32_t* inv_pos = field<uint32_t>(st, L, SSB_F_INV_POS);
    for (uint32_t i = tid; i < R; i += blockDim.x) {
        keys_sorted[i] = s_keys[i];
        point_list[i] = s_vals[i] & 1023u;
        inv_pos[s_vals[i] >> 10] = i;
    }
    // ---- ranges: dense (identifyTileRanges + memset) and the compact active-tile list
    uint2* ranges = field<uint2>(st, L, SSB_F_RANGES);
    const int tiles_total = vi.gx * vi.gy;
    for (int i = tid; i < tiles_total; i += blockDim.x) ranges[i] = make_uint2(0u, 0u);
    __syncthreads();
    uint32_t* tile_ids = field<uint32_t>(st, L, SSB_F_TILE_IDS);
    uint2* tile_ranges = field<uint2>(st, L, SSB_F_TILE_RANGES);
    if (tid < 32) {   // warp 0: ordered compaction of run starts
        uint32_t nact = 0;
        for (uint32_t base = 0; base < R; base += 32) {
            const uint32_t i = base + tid;
            bool start = false;
            uint32_t tile = 0;
            if (i < R) {
                tile = (uint32_t)(s_keys[i] >> 32);
                start = (i == 0) || ((uint32_t)(s_keys[i - 1] >> 32) != tile);
            }
            const uint32_t m = __ballot_sync(0xFFFFFFFFu, start);
            if (start) {
                const uint32_t a = nact + __popc(m & ((1u << tid) - 1u));
                // find the end of this run
                uint32_t e = i + 1;
                while (e < R && (uint32_t)(s_keys[e] >> 32) == tile) e++;
                tile_ids[a] = tile;
                tile_ranges[a] = make_uint2(i, e);
                ranges[tile] = make_uint2(i, e);
            }
            nact += __popc(m);
        }
        if (tid == 0) { s_nact = (int)nact; s_R = (int)R; }
    }
    __syncthreads();
    if (tid == 0) {
        int* hdr = field<int>(st, L, SSB_F_HEADER);
        hdr[0] = s_R; hdr[1] = s_nact; hdr[2] = (R_full > (uint32_t)rcap) ? (int)SSB_STATUS_R_OVERFLOW : 0;
        hdr[3] = P; hdr[4] = vi.W; hdr[5] = vi.H; hdr[6] = rcap; hdr[7] = (int)R_full;
    }
}

// ------------------------------------------------------------------------------------------ render fwd
// The dense contract says "a freshly written [C,H,W] image", and >97 % of it is zeros.  Two kernels:
//   fill_zero_kernel      pure streaming 256-bit stores over each view's image + inverse depth (HBM-write bound);
//   render_active_kernel  CTAs loop over the view's ACTIVE tiles only (compact list from bin_kernel) and overwrite
//                         them with the composited values (~3 % of the bytes are written twice).
// The reference writes every element twice (torch::full, then renderCUDA on all tiles) plus 12 B/pixel of side
// buffers (final_T, n_contrib, ranges sized W*H); here backward recomputes those on the active tiles instead.
constexpr int FILL_THREADS = 512;

// 256-bit global store (STG.E.256, new on sm_100): measured 7.23 TB/s for a pure fill at 148x64 CTAs x 512 threads
// vs 5.99 TB/s for the 128-bit version at 148x8 x 256 (scripts/fillbench.cu; cudaMemsetAsync reaches 7.18 TB/s).
__device__ __forceinline__ void store_zero_256(float* q, float z) {
    asm volatile("st.global.v8.f32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" :: "l"(q), "f"(z) : "memory");
}

// z = 0, or NaN for a view whose (Gaussian,tile) pairs outgrew r_capacity: its tile lists are truncated, so instead of a
// plausible-looking wrong image the caller gets one that poisons every reduction over it (no host sync needed to notice).
__device__ __forceinline__ void fill_zero(float* __restrict__ p, size_t n, size_t g, size_t stride, float z) {
    size_t head = ((32 - (reinterpret_cast<uintptr_t>(p) & 31)) & 31) >> 2;      // floats before 32-B alignment
    if (head > n) head = n;
    if (g < head) p[g] = z;
    float* p8 = p + head;
    const size_t n8 = (n - head) >> 3;
    size_t i = g;
    for (; i + stride < n8; i += 2 * stride) {      // 2 independent 32-B stores in flight per thread
        store_zero_256(p8 + 8 * i, z);
        store_zero_256(p8 + 8 * (i + stride), z);
    }
    for (; i < n8; i += stride) store_zero_256(p8 + 8 * i, z);
    const size_t tail = head + (n8 << 3) + g;       // < 8 floats left
    if (g < 8 && tail < n) p[tail] = z;
}

__global__ void __launch_bounds__(FILL_THREADS)
fill_zero_kernel(ssb_cameras cams, int C, StateLayout L, const char* __restrict__ state, float* __restrict__ out_color,
                 const int64_t* __restrict__ color_offsets, float* __restrict__ out_invdepth, const int64_t* __restrict__ invdepth_offsets)
{
    const int b = blockIdx.y;
    const float z = cfield<int>(state + (size_t)b * L.total, L, SSB_F_HEADER)[2] ? __int_as_float(0x7fc00000) : 0.0f;
    const int cam = b % cams.n_views;
    const int W = cams.dims ? cams.dims[2 * cam] : cams.W0;
    const int H = cams.dims ? cams.dims[2 * cam + 1] : cams.H0;
    const size_t HW = (size_t)H * W;
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    fill_zero(out_color + (color_offsets ? color_offsets[b] : (int64_t)b * C * (int64_t)cams.H0 * cams.W0), (size_t)C * HW, g, stride, z);
    if (out_invdepth)
        fill_zero(out_invdepth + (invdepth_offsets ? invdepth_offsets[b] : (int64_t)b * (int64_t)cams.H0 * cams.W0), HW, g, stride, z);
}

template <int C>
__global__ void __launch_bounds__(TILE * TILE)
render_active_kernel(ssb_gaussians g, ssb_cameras cams, StateLayout L, const char* __restrict__ state,
                     float* __restrict__ out_color, const int64_t* __restrict__ color_offsets,
                     float* __restrict__ out_invdepth, const int64_t* __restrict__ invdepth_offsets)
{
    const int b = blockIdx.y;
    const int frame = b / cams.n_views, cam = b % cams.n_views;
    const int W = cams.dims ? cams.dims[2 * cam] : cams.W0;
    const int H = cams.dims ? cams.dims[2 * cam + 1] : cams.H0;
    const int gx = (W + TILE - 1) / TILE;
    const char* st = state + (size_t)b * L.total;
    const int n_active = cfield<int>(st, L, SSB_F_HEADER)[1];
    const uint32_t* tile_ids = cfield<uint32_t>(st, L, SSB_F_TILE_IDS);
    const uint2* tile_ranges = cfield<uint2>(st, L, SSB_F_TILE_RANGES);
    const size_t HW = (size_t)H * W;
    float* color = out_color + (color_offsets ? color_offsets[b] : (int64_t)b * C * (int64_t)cams.H0 * cams.W0);
    float* invd = out_invdepth ? out_invdepth + (invdepth_offsets ? invdepth_offsets[b] : (int64_t)b * (int64_t)cams.H0 * cams.W0) : nullptr;
    const int tid = threadIdx.y * TILE + threadIdx.x;
    const uint32_t* point_list = cfield<uint32_t>(st, L, SSB_F_POINT_LIST);
    const float2* means2D = cfield<float2>(st, L, SSB_F_MEANS2D);
    const float4* conic_opacity = cfield<float4>(st, L, SSB_F_CONIC_OPACITY);
    const float* depths = cfield<float>(st, L, SSB_F_DEPTHS);
    const float* feats = g.features + (g.features_per_frame ? (size_t)frame * g.P * C : 0);
    __shared__ int s_id[TILE * TILE];
    __shared__ float2 s_xy[TILE * TILE];
    __shared__ float4 s_co[TILE * TILE];
    __shared__ float s_invd[TILE * TILE];

    for (int a = blockIdx.x; a < n_active; a += gridDim.x) {
        // ---- front-to-back compositing of one active tile (forward.cu:278-401)
        const uint32_t tile = tile_ids[a];
        const uint2 range = tile_ranges[a];
        const int px = (int)(tile % gx) * TILE + threadIdx.x, py = (int)(tile / gx) * TILE + threadIdx.y;
        const bool inside = px < W && py < H;
        const float pxf = (float)px, pyf = (float)py;
        bool done = !inside;
        float T = 1.0f, inv_acc = 0.0f;
        float acc[C];
#pragma unroll
        for (int c = 0; c < C; c++) acc[c] = 0.f;
        int todo = (int)(range.y - range.x);
        for (uint32_t base = range.x; base < range.y; base += TILE * TILE, todo -= TILE * TILE) {
            if (__syncthreads_count(done) == TILE * TILE) break;
            if (base + tid < range.y) {
                const int id = (int)point_list[base + tid];
                s_id[tid] = id;
                s_xy[tid] = means2D[id];
                s_co[tid] = conic_opacity[id];
                s_invd[tid] = __frcp_rn(depths[id]);
            }
            __syncthreads();
            const int n = min(TILE * TILE, todo);
            for (int j = 0; !done && j < n; j++) {
                const float2 xy = s_xy[j];
                const float4 co = s_co[j];
                float dx, dy, G, alpha;
                if (!pair_alpha(xy.x, xy.y, co.x, co.y, co.z, co.w, pxf, pyf, dx, dy, G, alpha)) continue;
                const float test_T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
                if (test_T < T_EPS) { done = true; continue; }
                const float* f = feats + (size_t)s_id[j] * C;
#pragma unroll
                for (int c = 0; c < C; c++) acc[c] = __fmaf_rn(T, __fmul_rn(alpha, __ldg(f + c)), acc[c]);
                inv_acc = __fmaf_rn(T, __fmul_rn(alpha, s_invd[j]), inv_acc);
                T = test_T;
            }
        }
        if (inside) {
            const size_t pix = (size_t)py * W + px;
#pragma unroll
            for (int c = 0; c < C; c++) color[(size_t)c * HW + pix] = acc[c];
            if (invd) invd[pix] = inv_acc;
        }
        __syncthreads();    // staging buffers are reused by the next tile
    }
}

// ------------------------------------------------------------------------------------------ render bwd
// NV = 7 + C partial sums per (tile, Gaussian) entry:
//   0,1 dL/dmean2D (x,y)   2,3,4 dL/dconic (x,y,w)   5 dL/dopacity   6 dL/dinvdepth   7.. dL/dfeatures
constexpr int BWD_CHUNK = 32;

template <int C>
__global__ void __launch_bounds__(TILE * TILE)
render_bwd_kernel(ssb_gaussians g, ssb_cameras cams, StateLayout L, const char* __restrict__ state,
                  const float* __restrict__ dL_dcolor, const int64_t* __restrict__ color_offsets,
                  const float* __restrict__ dL_dinvdepth, const int64_t* __restrict__ invdepth_offsets,
                  float* __restrict__ scratch, size_t scratch_stride)
{
    constexpr int NV = 7 + C;
    constexpr int NW = TILE * TILE / 32;
    const int b = blockIdx.y;
    const int frame = b / cams.n_views, cam = b % cams.n_views;
    const int W = cams.dims ? cams.dims[2 * cam] : cams.W0;
    const int H = cams.dims ? cams.dims[2 * cam + 1] : cams.H0;
    const int gx = (W + TILE - 1) / TILE;
    const size_t HW = (size_t)H * W;
    const char* st = state + (size_t)b * L.total;
    const int* hdr = cfield<int>(st, L, SSB_F_HEADER);
    const int n_active = hdr[1];
    const uint32_t* tile_ids = cfield<uint32_t>(st, L, SSB_F_TILE_IDS);
    const uint2* tile_ranges = cfield<uint2>(st, L, SSB_F_TILE_RANGES);
    const uint32_t* point_list = cfield<uint32_t>(st, L, SSB_F_POINT_LIST);
    const float2* means2D = cfield<float2>(st, L, SSB_F_MEANS2D);
    const float4* conic_opacity = cfield<float4>(st, L, SSB_F_CONIC_OPACITY);
    const float* depths = cfield<float>(st, L, SSB_F_DEPTHS);
    const float* feats = g.features + (g.features_per_frame ? (size_t)frame * g.P * C : 0);
    const float* dcol = dL_dcolor + (color_offsets ? color_offsets[b] : (int64_t)b * C * (int64_t)cams.H0 * cams.W0);
    const float* dinv = dL_dinvdepth ? dL_dinvdepth + (invdepth_offsets ? invdepth_offsets[b] : (int64_t)b * (int64_t)cams.H0 * cams.W0) : nullptr;
    float* part = scratch + (size_t)b * scratch_stride;

    __shared__ int s_id[BWD_CHUNK];
    __shared__ float2 s_xy[BWD_CHUNK];
    __shared__ float4 s_co[BWD_CHUNK];
    __shared__ float s_invd[BWD_CHUNK];
    __shared__ float s_feat[BWD_CHUNK][C];
    __shared__ float s_red[BWD_CHUNK][NW][NV];

    const int tid = threadIdx.y * TILE + threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;

    for (int a = blockIdx.x; a < n_active; a += gridDim.x) {
        const uint32_t tile = tile_ids[a];
        const uint2 range = tile_ranges[a];
        const int n_entries = (int)(range.y - range.x);
        const int px = (tile % gx) * TILE + threadIdx.x, py = (tile / gx) * TILE + threadIdx.y;
        const bool inside = px < W && py < H;
        const float pxf = (float)px, pyf = (float)py;
        const size_t pix = (size_t)py * W + px;

        // ---- pass 1: forward replay -> T_final, last contributor (what the reference stores per pixel)
        float T = 1.0f;
        uint32_t contributor = 0, last_contributor = 0;
        {
            bool done = !inside;
            for (int base = 0; base < n_entries; base += BWD_CHUNK) {
                __syncthreads();
                if (tid < BWD_CHUNK && base + tid < n_entries) {
                    const int id = (int)point_list[range.x + base + tid];
                    s_xy[tid] = means2D[id];
                    s_co[tid] = conic_opacity[id];
                }
                __syncthreads();
                const int n = min(BWD_CHUNK, n_entries - base);
                for (int j = 0; !done && j < n; j++) {
                    contributor++;
                    const float2 xy = s_xy[j];
                    const float4 co = s_co[j];
                    float dx, dy, G, alpha;
                    if (!pair_alpha(xy.x, xy.y, co.x, co.y, co.z, co.w, pxf, pyf, dx, dy, G, alpha)) continue;
                    const float test_T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
                    if (test_T < T_EPS) { done = true; continue; }
                    T = test_T;
                    last_contributor = contributor;
                }
            }
        }
        const float T_final = inside ? T : 0.f;
        if (!inside) last_contributor = 0;

        // ---- pass 2: back-to-front (backward.cu:452-638)
        float dpix[C];
#pragma unroll
        for (int c = 0; c < C; c++) dpix[c] = inside ? dcol[(size_t)c * HW + pix] : 0.f;
        const float dinv_pix = (inside && dinv) ? dinv[pix] : 0.f;
        float accum_rec[C], last_color[C];
#pragma unroll
        for (int c = 0; c < C; c++) { accum_rec[c] = 0.f; last_color[c] = 0.f; }
        float accum_invd_rec = 0.f, last_invd = 0.f, last_alpha = 0.f;
        T = T_final;
        contributor = (uint32_t)n_entries;
        for (int base = 0; base < n_entries; base += BWD_CHUNK) {
            const int n = min(BWD_CHUNK, n_entries - base);
            __syncthreads();
            if (tid < n) {   // reverse order: chunk slot j <-> sorted entry (range.y - 1 - base - j)
                const int id = (int)point_list[range.y - 1 - base - tid];
                s_id[tid] = id;
                s_xy[tid] = means2D[id];
                s_co[tid] = conic_opacity[id];
                s_invd[tid] = 1.f / depths[id];
            }
            __syncthreads();
            for (int i = tid; i < n * C; i += TILE * TILE) s_feat[i / C][i % C] = feats[(size_t)s_id[i / C] * C + (i % C)];
            __syncthreads();
            for (int j = 0; j < n; j++) {
                contributor--;
                float v[NV];
#pragma unroll
                for (int k = 0; k < NV; k++) v[k] = 0.f;
                bool active = false;
                if (inside && contributor < last_contributor) {
                    const float2 xy = s_xy[j];
                    const float4 co = s_co[j];
                    float dx, dy, G, alpha;
                    if (pair_alpha(xy.x, xy.y, co.x, co.y, co.z, co.w, pxf, pyf, dx, dy, G, alpha)) {
                        active = true;
                        T = T / (1.f - alpha);
                        const float dchannel_dcolor = alpha * T;
                        float dL_dalpha = 0.f;
#pragma unroll
                        for (int c = 0; c < C; c++) {
                            const float col = s_feat[j][c];
                            accum_rec[c] = last_alpha * last_color[c] + (1.f - last_alpha) * accum_rec[c];
                            last_color[c] = col;
                            dL_dalpha += (col - accum_rec[c]) * dpix[c];
                            v[7 + c] = dchannel_dcolor * dpix[c];
                        }
                        if (dinv) {
                            const float invd = s_invd[j];
                            accum_invd_rec = last_alpha * last_invd + (1.f - last_alpha) * accum_invd_rec;
                            last_invd = invd;
                            dL_dalpha += (invd - accum_invd_rec) * dinv_pix;
                            v[6] = dchannel_dcolor * dinv_pix;
                        }
                        dL_dalpha *= T;
                        last_alpha = alpha;
                        // background is defined as 0 (the reference reads past its 3-float bg)
                        const float dL_dG = co.w * dL_dalpha;
                        const float gdx = G * dx, gdy = G * dy;
                        const float dG_ddelx = -gdx * co.x - gdy * co.y;
                        const float dG_ddely = -gdy * co.z - gdx * co.y;
                        v[0] = dL_dG * dG_ddelx * ddelx_dx;
                        v[1] = dL_dG * dG_ddely * ddely_dy;
                        v[2] = -0.5f * gdx * dx * dL_dG;
                        v[3] = -0.5f * gdx * dy * dL_dG;
                        v[4] = -0.5f * gdy * dy * dL_dG;
                        v[5] = G * dL_dalpha;
                    }
                }
                // warp-level reduction by shuffles; skipped when no lane of the warp contributes
                if (__any_sync(0xFFFFFFFFu, active)) {
#pragma unroll
                    for (int k = 0; k < NV; k++) {
                        float x = v[k];
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xFFFFFFFFu, x, o);
                        if (lane == 0) s_red[j][warp][k] = x;
                    }
                } else if (lane < NV) {
                    s_red[j][warp][lane] = 0.f;
                }
            }
            __syncthreads();
            // fixed-order sum over the 8 warps -> scratch[sorted position][NV]
            for (int i = tid; i < n * NV; i += TILE * TILE) {
                const int j = i / NV, k = i - j * NV;
                float x = 0.f;
#pragma unroll
                for (int w = 0; w < NW; w++) x += s_red[j][w][k];
                part[(size_t)(range.y - 1 - base - j) * NV + k] = x;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------ per-Gaussian bwd
template <int C>
__global__ void __launch_bounds__(128)
gauss_bwd_kernel(ssb_gaussians g, ssb_cameras cams, StateLayout L, const char* __restrict__ state,
                 const float* __restrict__ scratch, size_t scratch_stride, bool has_invdepth,
                 float* __restrict__ dL_dmeans3D, float* __restrict__ dL_dmeans2D, float* __restrict__ dL_dscales,
                 float* __restrict__ dL_drotations, float* __restrict__ dL_dopacity, float* __restrict__ dL_dfeatures,
                 float* __restrict__ dL_dcov3D, float* __restrict__ dL_dconic)
{
    constexpr int NV = 7 + C;
    const int b = blockIdx.x;
    const int frame = b / cams.n_views, cam = b % cams.n_views;
    __shared__ float s_view[16], s_proj[16];
    if (threadIdx.x < 16) s_view[threadIdx.x] = cams.viewmatrix[16 * cam + threadIdx.x];
    else if (threadIdx.x < 32) s_proj[threadIdx.x - 16] = cams.projmatrix[16 * cam + threadIdx.x - 16];
    __syncthreads();
    const ViewInfo vi = view_info(cams, cam);
    const char* st = state + (size_t)b * L.total;
    const int* hdr = cfield<int>(st, L, SSB_F_HEADER);
    const uint32_t R = (uint32_t)hdr[0];
    const uint32_t* tiles_touched = cfield<uint32_t>(st, L, SSB_F_TILES_TOUCHED);
    const uint32_t* offsets = cfield<uint32_t>(st, L, SSB_F_POINT_OFFSETS);
    const uint32_t* inv_pos = cfield<uint32_t>(st, L, SSB_F_INV_POS);
    const float* cov3Ds = cfield<float>(st, L, SSB_F_COV3D);
    const float* part = scratch + (size_t)b * scratch_stride;
    const int P = g.P;
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        const size_t gi = (size_t)frame * P + i, oi = (size_t)b * P + i;
        float v[NV];
#pragma unroll
        for (int k = 0; k < NV; k++) v[k] = 0.f;
        const bool visible = tiles_touched[i] > 0;     // == radii > 0
        if (visible) {
            const uint32_t o0 = (i == 0) ? 0u : offsets[i - 1], o1 = offsets[i];
            for (uint32_t o = o0; o < o1 && o < R; o++) {   // emission order: deterministic
                const float* p = part + (size_t)inv_pos[o] * NV;
#pragma unroll
                for (int k = 0; k < NV; k++) v[k] += p[k];
            }
        }
        SplatGrad sg;
#pragma unroll
        for (int k = 0; k < 3; k++) { sg.dmean[k] = 0.f; sg.dscale[k] = 0.f; }
#pragma unroll
        for (int k = 0; k < 4; k++) sg.drot[k] = 0.f;
#pragma unroll
        for (int k = 0; k < 6; k++) sg.dcov[k] = 0.f;
        if (visible) {
            const bool has_sr = g.cov3D_precomp == nullptr;
            sg = gaussian_backward(g.means3D[3 * gi], g.means3D[3 * gi + 1], g.means3D[3 * gi + 2], cov3Ds + 6 * i, has_sr,
                                   has_sr ? g.scales[3 * gi] : 0.f, has_sr ? g.scales[3 * gi + 1] : 0.f, has_sr ? g.scales[3 * gi + 2] : 0.f,
                                   g.scale_modifier,
                                   has_sr ? g.rotations[4 * gi] : 1.f, has_sr ? g.rotations[4 * gi + 1] : 0.f,
                                   has_sr ? g.rotations[4 * gi + 2] : 0.f, has_sr ? g.rotations[4 * gi + 3] : 0.f,
                                   s_view, s_proj, vi.focal_x, vi.focal_y, vi.tan_fovx, vi.tan_fovy,
                                   v[0], v[1], v[2], v[3], v[4], v[6], has_invdepth);
        }
        if (dL_dmeans3D) { dL_dmeans3D[3 * oi] = sg.dmean[0]; dL_dmeans3D[3 * oi + 1] = sg.dmean[1]; dL_dmeans3D[3 * oi + 2] = sg.dmean[2]; }
        if (dL_dmeans2D) { dL_dmeans2D[3 * oi] = v[0]; dL_dmeans2D[3 * oi + 1] = v[1]; dL_dmeans2D[3 * oi + 2] = 0.f; }
        if (dL_dscales) { dL_dscales[3 * oi] = sg.dscale[0]; dL_dscales[3 * oi + 1] = sg.dscale[1]; dL_dscales[3 * oi + 2] = sg.dscale[2]; }
        if (dL_drotations) {
#pragma unroll
            for (int k = 0; k < 4; k++) dL_drotations[4 * oi + k] = sg.drot[k];
        }
        if (dL_dopacity) dL_dopacity[oi] = v[5];
        if (dL_dfeatures) {
#pragma unroll
            for (int c = 0; c < C; c++) dL_dfeatures[oi * C + c] = v[7 + c];
        }
        if (dL_dcov3D) {
#pragma unroll
            for (int k = 0; k < 6; k++) dL_dcov3D[6 * oi + k] = sg.dcov[k];
        }
        if (dL_dconic) { dL_dconic[4 * oi] = v[2]; dL_dconic[4 * oi + 1] = v[3]; dL_dconic[4 * oi + 2] = 0.f; dL_dconic[4 * oi + 3] = v[4]; }
    }
}

__global__ void mark_visible_kernel(int P, const float* __restrict__ means3D, const float* __restrict__ view, uint8_t* __restrict__ present) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const float depth = affine3(view[2], means3D[3 * i], view[6], means3D[3 * i + 1], view[10], means3D[3 * i + 2], view[14]);
    present[i] = depth > NEAR_Z;
}

// ------------------------------------------------------------------------------------------ host side
static int next_pow2(int x) { int n = 1; while (n < x) n <<= 1; return n; }

#define SSB_DISPATCH_C(Cval, ...)                                   \
    switch (Cval) {                                                 \
        case 1:  { constexpr int CC = 1;  __VA_ARGS__; } break;     \
        case 3:  { constexpr int CC = 3;  __VA_ARGS__; } break;     \
        case 15: { constexpr int CC = 15; __VA_ARGS__; } break;     \
        case 17: { constexpr int CC = 17; __VA_ARGS__; } break;     \
        case 19: { constexpr int CC = 19; __VA_ARGS__; } break;     \
        default: return SSB_ERR_UNSUPPORTED;                        \
    }

// W0 x H0 is the largest view of the rig (per-view sizes, if any, live in cams->dims on the device): it sizes the state.
static void max_dims(const ssb_cameras* cams, int& Wmax, int& Hmax) {
    Wmax = cams->W0; Hmax = cams->H0;
}

}  // namespace ssb

using namespace ssb;

extern "C" {

int ssb_channels_supported(int C) { return C == 1 || C == 3 || C == 15 || C == 17 || C == 19; }

size_t ssb_state_bytes(int P, int W, int H, int r_capacity) { return state_layout(P, W, H, r_capacity).total; }

int64_t ssb_state_field_offset(int P, int W, int H, int r_capacity, int f) {
    if (f < 0 || f >= SSB_F_COUNT) return -1;
    return (int64_t)state_layout(P, W, H, r_capacity).off[f];
}

size_t ssb_backward_scratch_bytes(int C, int r_capacity) { return align_up((size_t)r_capacity * (7 + C) * sizeof(float), 512); }

static int check_common(int n_frames, const ssb_gaussians* g, const ssb_cameras* cams, int rcap) {
    if (!g || !cams || n_frames < 0 || g->P < 0 || cams->n_views <= 0) return SSB_ERR_INVALID;
    if (!ssb_channels_supported(g->C)) return SSB_ERR_UNSUPPORTED;
    if (g->P > 1024) return SSB_ERR_CAPACITY;               // sort payload packs the Gaussian id in 10 bits
    if (rcap <= 0 || rcap > (1 << 14)) return SSB_ERR_CAPACITY;
    if (cams->W0 <= 0 || cams->H0 <= 0) return SSB_ERR_INVALID;  // W0/H0 must be the max over views (state sizing)
    if (g->P > 0 && (!g->means3D || !g->opacities || !g->features)) return SSB_ERR_INVALID;
    if (g->P > 0 && !g->cov3D_precomp && (!g->scales || !g->rotations)) return SSB_ERR_INVALID;
    return SSB_OK;
}

int ssb_rasterize_forward(int n_frames, const ssb_gaussians* g, const ssb_cameras* cams, int rcap,
                          float* out_color, const int64_t* color_offsets, float* out_invdepth,
                          const int64_t* invdepth_offsets, int* radii, void* state, void* stream_)
{
    int rc = check_common(n_frames, g, cams, rcap);
    if (rc != SSB_OK) return rc;
    if (!out_color || !state || (g->P > 0 && !radii)) return SSB_ERR_INVALID;
    cudaStream_t stream = (cudaStream_t)stream_;
    const int B = n_frames * cams->n_views;
    if (B == 0) return SSB_OK;
    int Wmax, Hmax;
    max_dims(cams, Wmax, Hmax);
    const StateLayout L = state_layout(g->P, Wmax, Hmax, rcap);
    const int n2 = next_pow2(rcap);
    const size_t smem = (size_t)n2 * 12 + (size_t)(g->P > 0 ? g->P : 1) * 4 + 16;
    if (smem > 48 * 1024) {
        if (cudaFuncSetAttribute(bin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return ssb_set_cuda_error(cudaGetLastError());
    }
    bin_kernel<<<B, BIN_THREADS, smem, stream>>>(*g, *cams, rcap, n2, L, Wmax, Hmax, (char*)state, radii);
    // zero fill: ~148x64 CTAs in total (the fill microbench keeps gaining up to there), but at least 2 stores per thread
    int fx = (148 * 64 + B - 1) / B;
    const long long work = ((long long)g->C * Wmax * Hmax / 8 + 2 * FILL_THREADS - 1) / (2 * FILL_THREADS);
    if (fx > work) fx = (int)work;
    fx = fx < 1 ? 1 : fx;
    fill_zero_kernel<<<dim3(fx, B), FILL_THREADS, 0, stream>>>(*cams, g->C, L, (const char*)state, out_color, color_offsets, out_invdepth,
                                                               invdepth_offsets);
    if (g->P > 0) {
        int G = (148 * 8 + B - 1) / B;
        G = G < 4 ? 4 : (G > 256 ? 256 : G);
        const dim3 grid(G, B), block(TILE, TILE);
        SSB_DISPATCH_C(g->C, render_active_kernel<CC><<<grid, block, 0, stream>>>(*g, *cams, L, (const char*)state, out_color,
                                                                                  color_offsets, out_invdepth, invdepth_offsets));
    }
    return ssb_set_cuda_error(cudaGetLastError());
}

int ssb_rasterize_backward(int n_frames, const ssb_gaussians* g, const ssb_cameras* cams, int rcap,
                           const float* dL_dcolor, const int64_t* color_offsets, const float* dL_dinvdepth,
                           const int64_t* invdepth_offsets, const void* state, void* scratch,
                           float* dL_dmeans3D, float* dL_dmeans2D, float* dL_dscales, float* dL_drotations,
                           float* dL_dopacity, float* dL_dfeatures, float* dL_dcov3D, float* dL_dconic, void* stream_)
{
    int rc = check_common(n_frames, g, cams, rcap);
    if (rc != SSB_OK) return rc;
    if (!dL_dcolor || !state || !scratch) return SSB_ERR_INVALID;
    cudaStream_t stream = (cudaStream_t)stream_;
    const int B = n_frames * cams->n_views;
    if (B == 0 || g->P == 0) return SSB_OK;
    int Wmax, Hmax;
    max_dims(cams, Wmax, Hmax);
    const StateLayout L = state_layout(g->P, Wmax, Hmax, rcap);
    const size_t stride = ssb_backward_scratch_bytes(g->C, rcap) / sizeof(float);
    // CTAs per view looping over that view's active tiles: enough to fill 148 SMs at small B
    int G = (148 * 8 + B - 1) / B;
    G = G < 4 ? 4 : (G > 256 ? 256 : G);
    const dim3 grid(G, B), block(TILE, TILE);
    SSB_DISPATCH_C(g->C, render_bwd_kernel<CC><<<grid, block, 0, stream>>>(*g, *cams, L, (const char*)state, dL_dcolor, color_offsets,
                                                                            dL_dinvdepth, invdepth_offsets, (float*)scratch, stride));
    SSB_DISPATCH_C(g->C, gauss_bwd_kernel<CC><<<B, 128, 0, stream>>>(*g, *cams, L, (const char*)state, (const float*)scratch, stride,
                                                                      dL_dinvdepth != nullptr, dL_dmeans3D, dL_dmeans2D, dL_dscales,
                                                                      dL_drotations, dL_dopacity, dL_dfeatures, dL_dcov3D, dL_dconic));
    return ssb_set_cuda_error(cudaGetLastError());
}

int ssb_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix, uint8_t* present, void* stream_) {
    (void)projmatrix;
    if (P < 0 || (P > 0 && (!means3D || !viewmatrix || !present))) return SSB_ERR_INVALID;
    if (P == 0) return SSB_OK;
    mark_visible_kernel<<<(P + 255) / 256, 256, 0, (cudaStream_t)stream_>>>(P, means3D, viewmatrix, present);
    return ssb_set_cuda_error(cudaGetLastError());
}

}  // extern "C"
