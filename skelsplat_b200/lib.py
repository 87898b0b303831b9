"""ctypes binding of libskelsplat_b200.so (the C-ABI declared in include/skelsplat_b200.h).

There is NO fallback: if the shared library is missing or a call fails, this module
raises.  The library is built in-tree by ``__graft_entry__.build()`` /
``python -m skelsplat_b200.build`` (nvcc, sm_100a).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SKELSPLAT_B200_LIB", os.path.join(_HERE, "_lib", "libskelsplat_b200.so"))   # override: tuning variants only

SSB_OK = 0

# enum ssb_state_field
(F_HEADER, F_DEPTHS, F_MEANS2D, F_CONIC_OPACITY, F_COV3D, F_TILES_TOUCHED, F_POINT_OFFSETS, F_RECTS,
 F_KEYS_UNSORTED, F_VALS_UNSORTED, F_KEYS_SORTED, F_POINT_LIST, F_INV_POS, F_TILE_IDS, F_TILE_RANGES,
 F_RANGES, F_COUNT) = range(17)

LOSS_L2_GAUSSIAN, LOSS_L1, LOSS_L1_GAUSSIAN = 0, 1, 2

c_f32p = C.c_void_p


class Gaussians(C.Structure):
    _fields_ = [("P", C.c_int), ("C", C.c_int), ("means3D", c_f32p), ("scales", c_f32p), ("rotations", c_f32p),
                ("cov3D_precomp", c_f32p), ("opacities", c_f32p), ("features", c_f32p),
                ("features_per_frame", C.c_int), ("scale_modifier", C.c_float)]


class Cameras(C.Structure):
    _fields_ = [("n_views", C.c_int), ("viewmatrix", c_f32p), ("projmatrix", c_f32p), ("dims", C.c_void_p),
                ("tanfov", c_f32p), ("W0", C.c_int), ("H0", C.c_int), ("tanfovx0", C.c_float),
                ("tanfovy0", C.c_float), ("antialiasing", C.c_int)]


class OptConfig(C.Structure):
    _fields_ = [("J", C.c_int), ("V", C.c_int), ("iterations", C.c_int), ("accumulation_steps", C.c_int),
                ("lambda_consistency", C.c_float), ("limb_pairs", C.c_int * 8),
                ("lr_scaling", C.c_float), ("lr_rotation", C.c_float), ("lr_opacity", C.c_float),
                ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
                ("r_capacity", C.c_int), ("antialiasing", C.c_int), ("max_unrolled_list", C.c_int),
                ("resident_record_slots", C.c_int)]


_lib = None

EXPORTS = [
    "ssb_version", "ssb_source_hash", "ssb_source_manifest", "ssb_struct_size", "ssb_error_string", "ssb_last_cuda_error", "ssb_channels_supported",
    "ssb_state_bytes", "ssb_rasterize_forward", "ssb_backward_scratch_bytes", "ssb_rasterize_backward",
    "ssb_mark_visible", "ssb_state_field_offset",
    "ssb_loss_forward", "ssb_loss_backward", "ssb_limb_consistency", "ssb_adam_frame_step",
    "ssb_fused_ssim_forward", "ssb_fused_ssim_backward",
    "ssb_fused_ssim_mean_workspace_bytes", "ssb_fused_ssim_mean_forward", "ssb_fused_ssim_mean_backward",
    "ssb_optimize_workspace_bytes", "ssb_optimize_frames", "ssb_optimize_frames_debug",
    "ssb_triangulate_dlt", "ssb_heatmap_roi_rects", "ssb_heatmap_roi_offsets", "ssb_heatmap_roi_fill",
]


class SkelSplatLibraryError(RuntimeError):
    pass


def lib():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SkelSplatLibraryError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(skelsplat_b200 has no CPU or PyTorch fallback)")
        L = C.CDLL(LIB_PATH)
        missing = [n for n in EXPORTS if not hasattr(L, n)]
        if missing:
            raise SkelSplatLibraryError(f"{LIB_PATH} is stale, missing symbols {missing}: rebuild it")
        L.ssb_error_string.restype = C.c_char_p
        L.ssb_source_hash.restype = C.c_char_p
        L.ssb_source_manifest.restype = C.c_char_p
        if "SKELSPLAT_B200_LIB" not in os.environ:                  # tuning variants are built with other defines on purpose
            from . import build as _b
            if os.path.isdir(_b.CSRC) and _b.sources():
                built, now = L.ssb_source_hash().decode(), _b.source_hash()
                if built != now:
                    raise SkelSplatLibraryError(f"{LIB_PATH} is stale: built from sources {built}, the tree holds {now}: rebuild it "
                                                "(python -m skelsplat_b200.build)")
        L.ssb_last_cuda_error.restype = C.c_char_p
        L.ssb_state_bytes.restype = C.c_size_t
        L.ssb_backward_scratch_bytes.restype = C.c_size_t
        L.ssb_state_field_offset.restype = C.c_int64
        L.ssb_optimize_workspace_bytes.restype = C.c_size_t
        L.ssb_fused_ssim_mean_workspace_bytes.restype = C.c_size_t
        for which, struct in enumerate((Gaussians, Cameras, OptConfig)):      # the ctypes mirrors must match the header
            if L.ssb_struct_size(C.c_int(which)) != C.sizeof(struct):
                raise SkelSplatLibraryError(f"{LIB_PATH}: struct layout mismatch for {struct.__name__} "
                                            f"({L.ssb_struct_size(C.c_int(which))} vs {C.sizeof(struct)} bytes): rebuild the library")
        _lib = L
    return _lib


def check(rc, what):
    if rc != SSB_OK:
        L = lib()
        msg = L.ssb_error_string(C.c_int(rc)).decode()
        if rc == -3:
            msg += ": " + L.ssb_last_cuda_error().decode()
        raise SkelSplatLibraryError(f"{what} failed: {msg} (code {rc})")


def ptr(t):
    """Device pointer of a torch tensor (or NULL for None / empty)."""
    if t is None or t.numel() == 0:
        return C.c_void_p(0)
    return C.c_void_p(t.data_ptr())


def current_stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def state_bytes(P, W, H, rcap):
    return int(lib().ssb_state_bytes(C.c_int(P), C.c_int(W), C.c_int(H), C.c_int(rcap)))


def state_field_offset(P, W, H, rcap, field):
    return int(lib().ssb_state_field_offset(C.c_int(P), C.c_int(W), C.c_int(H), C.c_int(rcap), C.c_int(field)))


def backward_scratch_bytes(Cch, rcap):
    return int(lib().ssb_backward_scratch_bytes(C.c_int(Cch), C.c_int(rcap)))


def source_manifest():
    """{file: hash} of the sources the LOADED library was compiled from."""
    m = lib().ssb_source_manifest().decode()
    return dict(e.split(":") for e in m.split(",") if ":" in e)
