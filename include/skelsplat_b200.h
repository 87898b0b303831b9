/*
 * skelsplat_b200 -- C ABI of the B200-native (sm_100a) SkelSplat hot path.
 *
 * Plain pointers and sizes only: no torch types cross this boundary.  All pointers
 * are DEVICE pointers unless the name ends in _host.  Every entry point enqueues its
 * kernels on `stream` (a cudaStream_t passed as void*), never synchronises, and
 * returns an int status (SSB_OK or a negative SSB_ERR_*).  The library is stateless.
 *
 * Each entry point cites the reference interface it replaces (paths relative to the
 * reference repo; RAST = submodules/diff-gaussian-rasterization-{h36m,panoptic,op}).
 * The reference-side bindings a maintainer would add are shown in INTEGRATION.md.
 */
#ifndef SKELSPLAT_B200_H
#define SKELSPLAT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSB_OK               0
#define SSB_ERR_INVALID     -1   /* bad argument (null pointer, unsupported channel count, ...) */
#define SSB_ERR_CAPACITY    -2   /* a caller-provided buffer/capacity is too small */
#define SSB_ERR_CUDA        -3   /* a CUDA runtime call failed; see ssb_last_cuda_error() */
#define SSB_ERR_UNSUPPORTED -4

/* Device-side status bits written into a view's state header (see ssb_state_header). */
#define SSB_STATUS_R_OVERFLOW 1u /* more (Gaussian,tile) pairs than r_capacity: result invalid, re-run with more */
#define SSB_STATUS_ROI_OVERFLOW 2u /* heatmap patches need more than the roi_data capacity given: patches were skipped */
#define SSB_STATUS_ROI_TOO_WIDE 4u /* a heatmap patch is wider than 256 px (sigma > 31 px): outside the supported regime */

int         ssb_version(void);              /* 200: + ssb_optimize_frames_debug, ssb_source_hash, fp64 heatmap sigmas */
/* Hex digest of the kernel sources the library was compiled from (skelsplat_b200/build.py passes -DSSB_SOURCE_HASH):
 * a binding compares it with the sources it sits next to and refuses a stale build. */
const char* ssb_source_hash(void);
/* "file:hash,file:hash,...": one entry per kernel source / header the library was compiled from. */
const char* ssb_source_manifest(void);
int         ssb_struct_size(int which);     /* sizeof: 0 ssb_gaussians, 1 ssb_cameras, 2 ssb_opt_config; -1 otherwise */
const char* ssb_error_string(int code);
const char* ssb_last_cuda_error(void);
/* Channel counts the templated kernels are instantiated for (reference: NUM_CHANNELS in
 * RAST/cuda_rasterizer/config.h:15 = 17 / 19 / 15; 3 and 1 are kept for tests). */
int         ssb_channels_supported(int C);

/* ------------------------------------------------------------------------------------------
 * Dense-contract rasteriser, batched over B = n_frames * n_views views.
 * View b uses the Gaussians of frame (b / n_views) and camera (b % n_views).
 * Replaces: _C.rasterize_gaussians / RasterizeGaussiansCUDA      RAST/rasterize_points.cu:35-124
 *           CudaRasterizer::Rasterizer::forward                   RAST/cuda_rasterizer/rasterizer_impl.cu:198-341
 * ------------------------------------------------------------------------------------------ */

typedef struct ssb_gaussians {
    int          P;              /* Gaussians per frame */
    int          C;              /* feature channels (NUM_CHANNELS) */
    const float* means3D;        /* [F,P,3] */
    const float* scales;         /* [F,P,3] activated (exp applied) or NULL when cov3D_precomp is given */
    const float* rotations;      /* [F,P,4] (r,x,y,z), used un-normalised, or NULL */
    const float* cov3D_precomp;  /* [F,P,6] or NULL */
    const float* opacities;      /* [F,P] */
    const float* features;       /* [F,P,C], or [P,C] shared by all frames when features_per_frame == 0 */
    int          features_per_frame;
    float        scale_modifier;
} ssb_gaussians;

typedef struct ssb_cameras {
    int          n_views;        /* V; cameras are shared by all frames */
    const float* viewmatrix;     /* [V,16]  world_view_transform, torch row-major memory (= column-major W2C) */
    const float* projmatrix;     /* [V,16]  full_proj_transform */
    const int*   dims;           /* [V,2] (W,H) per view, or NULL => every view is W0 x H0 */
    const float* tanfov;         /* [V,2] (tanfovx,tanfovy) or NULL => tanfovx0/tanfovy0 */
    int          W0, H0;         /* image size when dims == NULL; otherwise the MAXIMUM over the views (sizes the state) */
    float        tanfovx0, tanfovy0;
    int          antialiasing;   /* 0 in every shipped config */
} ssb_cameras;

/* Bytes of opaque per-view state for P Gaussians, a W x H image (max over views) and at most
 * r_capacity (Gaussian,tile) pairs.  View b's state lives at state + b * ssb_state_bytes(). */
size_t ssb_state_bytes(int P, int W, int H, int r_capacity);

/* Forward.  out_color: view b's [C,H_b,W_b] image starts at out_color + color_offsets[b] floats
 * (color_offsets == NULL: b * C*H0*W0); out_invdepth likewise with invdepth_offsets (NULL: b*H0*W0)
 * and may be NULL (not rendered).  radii: [B,P] int32.  Every element of the outputs is written
 * (the reference's torch::full zero-fill is folded into the kernel). */
int ssb_rasterize_forward(int n_frames, const ssb_gaussians* g, const ssb_cameras* cams, int r_capacity,
                          float* out_color, const int64_t* color_offsets,
                          float* out_invdepth, const int64_t* invdepth_offsets,
                          int* radii, void* state, void* stream);

/* Backward.  Replaces _C.rasterize_gaussians_backward / RasterizeGaussiansBackwardCUDA
 * (RAST/rasterize_points.cu:126-223) and Rasterizer::backward (rasterizer_impl.cu:345-450).
 * dL_dcolor uses the same offsets as out_color; dL_dinvdepth may be NULL.
 * scratch: ssb_backward_scratch_bytes() bytes per view, contiguous for B views.
 * Per-view gradient outputs (any may be NULL): dL_dmeans3D [B,P,3], dL_dmeans2D [B,P,3],
 * dL_dscales [B,P,3], dL_drotations [B,P,4], dL_dopacity [B,P], dL_dfeatures [B,P,C],
 * dL_dcov3D [B,P,6], dL_dconic [B,P,4].  Deterministic: no atomics. */
size_t ssb_backward_scratch_bytes(int C, int r_capacity);
int ssb_rasterize_backward(int n_frames, const ssb_gaussians* g, const ssb_cameras* cams, int r_capacity,
                           const float* dL_dcolor, const int64_t* color_offsets,
                           const float* dL_dinvdepth, const int64_t* invdepth_offsets,
                           const void* state, void* scratch,
                           float* dL_dmeans3D, float* dL_dmeans2D, float* dL_dscales, float* dL_drotations,
                           float* dL_dopacity, float* dL_dfeatures, float* dL_dcov3D, float* dL_dconic,
                           void* stream);

/* Replaces _C.mark_visible / markVisible (RAST/rasterize_points.cu:225-244,
 * rasterizer_impl.cu:54-66,141-153): present[i] = view-space z > 0.2. */
int ssb_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                     uint8_t* present, void* stream);

/* Debug accessors for the bit-exact stage tests: byte offsets of the fields of
 * one view's state.  Field ids: */
enum ssb_state_field {
    SSB_F_HEADER = 0,      /* int32[8]: R, n_active_tiles, status, P, W, H, r_capacity, R_uncapped */
    SSB_F_DEPTHS,          /* f32[P] */
    SSB_F_MEANS2D,         /* f32[P,2] */
    SSB_F_CONIC_OPACITY,   /* f32[P,4] */
    SSB_F_COV3D,           /* f32[P,6] */
    SSB_F_TILES_TOUCHED,   /* u32[P] */
    SSB_F_POINT_OFFSETS,   /* u32[P] inclusive scan */
    SSB_F_RECTS,           /* u32[P,4] x0,y0,x1,y1 */
    SSB_F_KEYS_UNSORTED,   /* u64[r_capacity] */
    SSB_F_VALS_UNSORTED,   /* u32[r_capacity] */
    SSB_F_KEYS_SORTED,     /* u64[r_capacity] */
    SSB_F_POINT_LIST,      /* u32[r_capacity] sorted Gaussian ids */
    SSB_F_INV_POS,         /* u32[r_capacity] sorted position of each emission index */
    SSB_F_TILE_IDS,        /* u32[r_capacity] active tile ids */
    SSB_F_TILE_RANGES,     /* u32[r_capacity,2] (start,end) per active tile */
    SSB_F_RANGES,          /* u32[tiles,2] dense ranges, (0,0) for untouched tiles */
    SSB_F_COUNT
};
int64_t ssb_state_field_offset(int P, int W, int H, int r_capacity, int field);

/* ------------------------------------------------------------------------------------------
 * Fused dense losses on [C,H,W] images (reference: utils/loss_utils.py, ~10 ATen passes each).
 * kind: 0 = l2_gaussian (masked MSE, :86-100), 1 = l1 (:67-73), 2 = l1_gaussian / l1_masked (:103-117,173-192)
 * ------------------------------------------------------------------------------------------ */
#define SSB_LOSS_L2_GAUSSIAN 0
#define SSB_LOSS_L1          1
#define SSB_LOSS_L1_GAUSSIAN 2
/* Forward: sums[0] = sum of per-element loss over the mask, sums[1] = mask count (as double).
 * `sums` (double[2]) must be zeroed by the caller.  error_out (optional, [n]) receives the
 * dense per-element error the reference's l2_loss_gaussian also returns. */
int ssb_loss_forward(int kind, int64_t n, const float* render, const float* gt, double* sums,
                     float* error_out, void* stream);
/* Backward: grad[i] = dloss/drender[i] * (*grad_out) for reduction='mean' (scale = 1/count read
 * from sums[1] on the device; kind L1 uses 1/n).  The clamp(0,1) of the reference's render_*
 * (gaussian_renderer/__init__.py:129) stays a separate op of the caller. */
int ssb_loss_backward(int kind, int64_t n, const float* render, const float* gt, const double* sums,
                      const float* grad_out, float* grad, void* stream);

/* limb_3d_consistency_loss (utils/loss_utils.py:226-250) forward + gradient for F frames.
 * xyz [F,J,3]; pairs_host: 8 ints (l_arm a,b; r_arm a,b; l_leg a,b; r_leg a,b);
 * loss [F]; grad [F,J,3] is OVERWRITTEN with d loss / d xyz (unscaled). */
int ssb_limb_consistency(int n_frames, int J, const float* xyz, const int* pairs_host, float* loss, float* grad,
                         void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused SSIM.  Replaces fused_ssim_cuda.fusedssim / fusedssim_backward
 * (submodules/fused-ssim/ssim.cu:368-444, ssim.h:7-26): 11-tap sigma=1.5 Gaussian window, zero
 * padding.  Tensors are [B,CH,H,W].  dm_* may be NULL when train == 0.
 * Limits (SSB_ERR_CAPACITY beyond them): B*CH <= 65535 (grid.z), H*W < 2^31 (32-bit index math inside a plane).
 * ------------------------------------------------------------------------------------------ */
int ssb_fused_ssim_forward(int B, int CH, int H, int W, float C1, float C2, const float* img1, const float* img2,
                           float* ssim_map, float* dm_dmu1, float* dm_dsigma1_sq, float* dm_dsigma12, void* stream);
int ssb_fused_ssim_backward(int B, int CH, int H, int W, float C1, float C2, const float* img1, const float* img2,
                            const float* dL_dmap, const float* dm_dmu1, const float* dm_dsigma1_sq,
                            const float* dm_dsigma12, float* dL_dimg1, void* stream);

/* One torch.optim.Adam step (default foreach path, eps as given) for the four parameter tensors of ONE frame, written so that the
 * launch can be captured in a CUDA graph: the step-dependent host scalars of every step are precomputed into step_table (device,
 * [n_steps][5] floats: -lr_xyz/bc1, -lr_scaling/bc1, -lr_rotation/bc1, -lr_opacity/bc1, sqrt(bc2) with bc = 1 - beta^step, computed in
 * fp64 like torch's python floats and rounded to fp32) and the step index lives in *step_counter (device; advanced by the kernel).
 * The xyz gradient is the mean over the V slots of accumulated_grads [V,J,3] (train.py:215-218); exp_avg / exp_avg_sq are [11 J]
 * (xyz 3J | scaling 3J | rotation 4J | opacity J); one_minus_beta = fp32(1 - beta) with the subtraction done in fp64, as torch forms
 * the lerp / addcmul scalars.  Replaces scene/gaussian_model.py:217-218 + train.py:215-222 in the graphed
 * drop-in loop (skelsplat_b200/training.py:GraphedFrameOptimizer). */
int ssb_adam_frame_step(int J, int V, float* xyz, float* scaling, float* rotation, float* opacity, const float* accumulated_grads,
                        const float* g_scaling, const float* g_rotation, const float* g_opacity, float* exp_avg, float* exp_avg_sq,
                        const float* step_table, int n_steps, int* step_counter, float one_minus_beta1, float beta2, float one_minus_beta2,
                        float eps, void* stream);

/* The form fused_ssim() actually consumes -- map.mean() (fused_ssim/__init__.py:34-41) -- without materialising the map or
 * dL/dmap: forward writes the mean over the pixels at least `crop` away from the border (0: padding "same"; 5: "valid") to
 * *mean_out (device) and, when the dm_* pointers are given, the three derivative maps; backward takes dL/dmean as a device
 * scalar.  workspace: ssb_fused_ssim_mean_workspace_bytes() bytes (per-warp partial sums, summed in a fixed order). */
size_t ssb_fused_ssim_mean_workspace_bytes(int B, int CH, int H, int W);
int ssb_fused_ssim_mean_forward(int B, int CH, int H, int W, float C1, float C2, const float* img1, const float* img2, int crop,
                                float* mean_out, float* dm_dmu1, float* dm_dsigma1_sq, float* dm_dsigma12, void* workspace, void* stream);
int ssb_fused_ssim_mean_backward(int B, int CH, int H, int W, const float* img1, const float* img2, const float* grad_mean, int crop,
                                 const float* dm_dmu1, const float* dm_dsigma1_sq, const float* dm_dsigma12, float* dL_dimg1, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused per-frame optimiser: the whole train.py:130-233 iteration loop (render one view,
 * l2_gaussian + limb consistency, backward, gradient bookkeeping, Adam every accumulation_steps)
 * for many independent frames in one persistent kernel.  One CTA owns one frame.
 * ------------------------------------------------------------------------------------------ */
typedef struct ssb_opt_config {
    int   J;                    /* joints == Gaussians == channels (15 / 17 / 19) */
    int   V;                    /* views */
    int   iterations;           /* 500 */
    int   accumulation_steps;   /* 4 */
    float lambda_consistency;   /* 1e-5 */
    int   limb_pairs[8];
    float lr_scaling, lr_rotation, lr_opacity;
    float beta1, beta2, eps;    /* 0.9, 0.999, 1e-15 */
    int   r_capacity;           /* max (Gaussian,tile) pairs per view: a multiple of 32, 32..1024 */
    int   antialiasing;
    int   max_unrolled_list;    /* tuning knob (same arithmetic, different fp32 summation order): tile lists of 5 Gaussians take the unrolled
                                   register-resident path (0 = default = 5) or the chunked generic path (4) */
    int   resident_record_slots;/* tuning knob (bit-identical results): slots of a step group whose per-(tile,Gaussian) records are resident
                                   in shared memory at once.  0 = default = accumulation_steps (all of them; two 512-thread CTAs per SM when
                                   they fit, else one 1024-thread CTA); accumulation_steps/2 (4 slots only): the tile phase runs twice per
                                   Adam step over half the record storage, which keeps two CTAs per SM up to r_capacity 1024 -- measured
                                   slower on B200 for every shipped shape, kept as a knob. */
} ssb_opt_config;

/* lr_xyz_host: [iterations+1] learning rate of the xyz group at iteration i (host-computed in fp64
 * exactly as get_expon_lr_func, utils/general_utils.py:38-71, then rounded as torch does).
 * Per-frame state (updated in place): xyz [F,J,3], scaling_raw [F,J,3], rotation_raw [F,J,4],
 * opacity_raw [F,J].  Cameras as ssb_cameras.  GT heatmaps as FACTORED ROIs: roi_rect [F,V,J,4] int32
 * (x0,y0,w,h) and roi_offset [F,V,J] int64 into roi_data (float), where a patch is h + w floats
 * col[h] | row[w] and the heatmap value at window pixel (a,b) is the fp32 product col[a]*row[b]
 * (0 outside the window); the patches of one frame must be contiguous in roi_data.  Outputs:
 * final_loss [F] (loss of the last iteration) or NULL.  workspace: ssb_optimize_workspace_bytes(). */
size_t ssb_optimize_workspace_bytes(const ssb_opt_config* cfg, int n_frames);
int ssb_optimize_frames(const ssb_opt_config* cfg, int n_frames, const ssb_cameras* cams,
                        const double* lr_xyz_host,
                        float* xyz, float* scaling_raw, float* rotation_raw, float* opacity_raw,
                        const int* roi_rect, const int64_t* roi_offset, const float* roi_data,
                        float* final_loss, void* workspace, void* stream);

/* Debug accessor for the bit-exact binning tests: the same launch, and additionally the binning state of
 * frame dbg_frame at Adam step dbg_step is written to dbg_out (device, int32).  Per slot k (= iteration k of the step group)
 * 4 + 4*r_capacity words:  R, active tiles, view, status | point_list[r_capacity] (sorted Gaussian ids, the reference's
 * binningState.point_list, rasterizer_impl.cu:303-311) | inv_pos[r_capacity] (sorted position of each emission index) |
 * tile[r_capacity] ((y << 8) | x of each active tile, row-major order) | start[r_capacity] (first list entry of each active
 * tile = ranges[tile].x, rasterizer_impl.cu:116-138); unused entries are -1. */
int ssb_optimize_frames_debug(const ssb_opt_config* cfg, int n_frames, const ssb_cameras* cams,
                              const double* lr_xyz_host,
                              float* xyz, float* scaling_raw, float* rotation_raw, float* opacity_raw,
                              const int* roi_rect, const int64_t* roi_offset, const float* roi_data,
                              float* final_loss, void* workspace, int dbg_frame, int dbg_step, int* dbg_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Per-frame setup on the GPU.
 * ------------------------------------------------------------------------------------------ */
/* Batched DLT triangulation.  Replaces triangulate_poses / triangulate_points_multi_camera (triangulation.py:122-150).
 * P [V,3,4] fp64 projection matrices K[R|t]; poses_2d [F,V,J,2] fp64; out_xyz [F,J,3] fp64 (dehomogenised). */
int ssb_triangulate_dlt(int n_frames, int V, int J, const double* P, const double* poses_2d, double* out_xyz, void* stream);

/* Pseudo-GT heatmaps as factored ROI patches (col[h] | row[w] per patch, see ssb_optimize_frames).  Replaces generate_heatmaps
 * (utils/general_utils.py:175-304).  Two steps so the caller can size the packed buffer:  (1) rectangles / sigmas / peak
 * positions / patch sizes (h + w floats) for every (frame,view,joint); (2) after an exclusive scan of roi_size into
 * roi_offset, fill the profiles.
 * xyz [F,J,3], scaling_raw [F,J,3] (log), rotation_raw [F,J,4], poses_2d [F,V,J,2] fp32;
 * roi_rect [F,V,J,4] int32 (x0,y0,w,h), roi_sigma [F,V,J,2] (sigma_y, sigma_x), roi_center [F,V,J,2] int32 (x,y),
 * roi_size / roi_offset [F,V,J] int64, roi_data fp32. */
int ssb_heatmap_roi_rects(int n_frames, int J, const ssb_cameras* cams, const float* xyz, const float* scaling_raw,
                          const float* rotation_raw, const float* poses_2d, float scaling_modifier,
                          int* roi_rect, float* roi_sigma, int* roi_center, int64_t* roi_size, void* stream);
/* Exclusive scan roi_size[n] -> roi_offset[n] on the device (n = F*V*J); *total (device pointer, may be NULL) receives the
 * packed size.  With it the detections -> ROIs -> optimiser pipeline needs no host round trip. */
int ssb_heatmap_roi_offsets(int64_t n, const int64_t* roi_size, int64_t* roi_offset, int64_t* total, void* stream);
/* capacity: number of floats roi_data can hold, or -1 for "exactly sized by the caller".  A patch that does not fit (or is
 * wider than 256 px) is skipped and SSB_STATUS_ROI_OVERFLOW / SSB_STATUS_ROI_TOO_WIDE is OR-ed into *status (device int,
 * may be NULL). */
int ssb_heatmap_roi_fill(int n_frames, int J, const ssb_cameras* cams, const int* roi_rect, const float* roi_sigma,
                         const int* roi_center, const int64_t* roi_offset, float* roi_data, int64_t capacity, int* status,
                         void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SKELSPLAT_B200_H */
