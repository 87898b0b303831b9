// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// Thin extern "C" shim around the UNMODIFIED reference rasteriser so that the
// reference's own CUDA kernels can be driven from Python (ctypes) without
// compiling any torch translation unit.  The reference sources are compiled
// where they lie under /root/reference (see oracle/Makefile); nothing of them
// is copied into this repository.  Output: oracle/_ref/libref_rast_<variant>.so
//
// What it wraps (reference file:line, RAST = submodules/diff-gaussian-rasterization-*):
//   CudaRasterizer::Rasterizer::forward    RAST/cuda_rasterizer/rasterizer_impl.cu:198-341
//   CudaRasterizer::Rasterizer::backward   RAST/cuda_rasterizer/rasterizer_impl.cu:345-450
//   CudaRasterizer::Rasterizer::markVisible RAST/cuda_rasterizer/rasterizer_impl.cu:141-153
//   {Geometry,Image,Binning}State::fromChunk  RAST/cuda_rasterizer/rasterizer_impl.cu:155-194
// The torch glue the reference normally uses (RAST/rasterize_points.cu:35-223) only
// allocates tensors and forwards pointers; the Python side of this shim
// (oracle/ref_rasterizer.py) restates that allocation logic.
//
// Only tests/, __graft_entry__.smoke() and bench.py's reference / cpu_baseline
// legs may load the library built from this file.

#include <cstdint>
#include <cstddef>
#include <functional>
#include <stdexcept>
#include <cuda_runtime.h>

#include "cuda_rasterizer/config.h"
#include "cuda_rasterizer/rasterizer.h"
#include "cuda_rasterizer/rasterizer_impl.h"

namespace {
struct Arena {
    char* base;
    size_t cap;
    size_t used;
};
std::function<char*(size_t)> arena_fn(Arena* a) {
    return [a](size_t n) -> char* {
        a->used = n;
        if (n > a->cap) throw std::runtime_error("ref_shim: scratch arena too small");
        return a->base;
    };
}
}  // namespace

extern "C" {

int ref_num_channels() { return NUM_CHANNELS; }

// Bytes the reference asks for each opaque state buffer.
size_t ref_required_geom(size_t P) { return CudaRasterizer::required<CudaRasterizer::GeometryState>(P); }
size_t ref_required_image(size_t N) { return CudaRasterizer::required<CudaRasterizer::ImageState>(N); }
size_t ref_required_binning(size_t R) { return CudaRasterizer::required<CudaRasterizer::BinningState>(R); }

// Field offsets (bytes from `base`) of the carved state, for the bit-exact stage tests.
// geom: depths, clamped, internal_radii, means2D, cov3D, conic_opacity, rgb, tiles_touched, scanning_space, point_offsets
void ref_geom_layout(char* base, size_t P, size_t* off) {
    char* p = base;
    auto g = CudaRasterizer::GeometryState::fromChunk(p, P);
    off[0] = (char*)g.depths - base;
    off[1] = (char*)g.clamped - base;
    off[2] = (char*)g.internal_radii - base;
    off[3] = (char*)g.means2D - base;
    off[4] = (char*)g.cov3D - base;
    off[5] = (char*)g.conic_opacity - base;
    off[6] = (char*)g.rgb - base;
    off[7] = (char*)g.tiles_touched - base;
    off[8] = (char*)g.scanning_space - base;
    off[9] = (char*)g.point_offsets - base;
}
// image: accum_alpha, n_contrib, ranges
void ref_image_layout(char* base, size_t N, size_t* off) {
    char* p = base;
    auto s = CudaRasterizer::ImageState::fromChunk(p, N);
    off[0] = (char*)s.accum_alpha - base;
    off[1] = (char*)s.n_contrib - base;
    off[2] = (char*)s.ranges - base;
}
// binning: point_list, point_list_unsorted, point_list_keys, point_list_keys_unsorted, list_sorting_space
void ref_binning_layout(char* base, size_t R, size_t* off) {
    char* p = base;
    auto b = CudaRasterizer::BinningState::fromChunk(p, R);
    off[0] = (char*)b.point_list - base;
    off[1] = (char*)b.point_list_unsorted - base;
    off[2] = (char*)b.point_list_keys - base;
    off[3] = (char*)b.point_list_keys_unsorted - base;
    off[4] = (char*)b.list_sorting_space - base;
}

// Returns num_rendered (>=0) or a negative error code. used[3] receives the bytes
// the reference requested from each arena.
int ref_forward(
    int P, int D, int M,
    const float* background, int W, int H,
    const float* means3D, const float* shs, const float* colors_precomp,
    const float* opacities, const float* scales, float scale_modifier,
    const float* rotations, const float* cov3D_precomp,
    const float* viewmatrix, const float* projmatrix, const float* campos,
    float tan_fovx, float tan_fovy, int prefiltered,
    float* out_color, float* out_invdepth, int antialiasing, int* radii, int debug,
    char* geom, size_t geom_cap, char* binning, size_t binning_cap, char* img, size_t img_cap,
    size_t* used)
{
    Arena ag{geom, geom_cap, 0}, ab{binning, binning_cap, 0}, ai{img, img_cap, 0};
    int R;
    try {
        R = CudaRasterizer::Rasterizer::forward(
            arena_fn(&ag), arena_fn(&ab), arena_fn(&ai),
            P, D, M, background, W, H, means3D, shs, colors_precomp, opacities, scales,
            scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, campos,
            tan_fovx, tan_fovy, prefiltered != 0, out_color, out_invdepth, antialiasing != 0,
            radii, debug != 0);
    } catch (const std::exception&) {
        return -1;
    }
    if (used) { used[0] = ag.used; used[1] = ab.used; used[2] = ai.used; }
    return R;
}

int ref_backward(
    int P, int D, int M, int R,
    const float* background, int W, int H,
    const float* means3D, const float* shs, const float* colors_precomp,
    const float* opacities, const float* scales, float scale_modifier,
    const float* rotations, const float* cov3D_precomp,
    const float* viewmatrix, const float* projmatrix, const float* campos,
    float tan_fovx, float tan_fovy, const int* radii,
    char* geom, char* binning, char* img,
    const float* dL_dpix, const float* dL_dinvdepth_pix,
    float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
    float* dL_dinvdepth, float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh,
    float* dL_dscale, float* dL_drot, int antialiasing, int debug)
{
    try {
        CudaRasterizer::Rasterizer::backward(
            P, D, M, R, background, W, H, means3D, shs, colors_precomp, opacities, scales,
            scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, campos,
            tan_fovx, tan_fovy, radii, geom, binning, img, dL_dpix, dL_dinvdepth_pix,
            dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor, dL_dinvdepth, dL_dmean3D,
            dL_dcov3D, dL_dsh, dL_dscale, dL_drot, antialiasing != 0, debug != 0);
    } catch (const std::exception&) {
        return -1;
    }
    return 0;
}

void ref_mark_visible(int P, float* means3D, float* viewmatrix, float* projmatrix, bool* present) {
    CudaRasterizer::Rasterizer::markVisible(P, means3D, viewmatrix, projmatrix, present);
}

}  // extern "C"
