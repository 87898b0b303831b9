"""TEST INFRASTRUCTURE ONLY -- drives the UNMODIFIED reference CUDA rasteriser.

oracle/_ref/libref_rast_<variant>.so is the reference's own forward.cu / backward.cu /
rasterizer_impl.cu compiled for sm_100a (oracle/Makefile) behind the extern "C" shim
oracle/ref_shim.cu.  This module restates the reference's torch glue, which only
allocates tensors and forwards pointers:
  * RasterizeGaussiansCUDA / ...BackwardCUDA    RAST/rasterize_points.cu:35-223
  * _RasterizeGaussians / Settings / Rasterizer RAST/diff_gaussian_rasterization_h36m/__init__.py:44-207
so that "the reference op" exists on the GPU box without compiling a torch TU.
All reference kernels run on the legacy default stream, as in the reference.
It needs a GPU at run time; /root/reference is NOT needed at run time.
"""
import ctypes as C
import os
from typing import NamedTuple

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}
VARIANT_CHANNELS = {"h36m": 17, "panoptic": 19, "op": 15}


def available(variant="h36m"):
    return os.path.exists(os.path.join(_HERE, "_ref", f"libref_rast_{variant}.so")) and torch.cuda.is_available()


def lib(variant):
    if variant not in _LIBS:
        L = C.CDLL(os.path.join(_HERE, "_ref", f"libref_rast_{variant}.so"))
        for f in ("ref_required_geom", "ref_required_image", "ref_required_binning"):
            getattr(L, f).restype = C.c_size_t
            getattr(L, f).argtypes = [C.c_size_t]
        L.ref_forward.restype = C.c_int
        L.ref_backward.restype = C.c_int
        L.ref_num_channels.restype = C.c_int
        _LIBS[variant] = L
    return _LIBS[variant]


def _ptr(t):
    if t is None or t.numel() == 0:
        return C.c_void_p(0)
    return C.c_void_p(t.data_ptr())


class RefState:
    """The three opaque byte buffers + R that the reference's ctx carries to backward."""
    def __init__(self, geom, binning, img, R, P, W, H, variant):
        self.geom, self.binning, self.img, self.R, self.P, self.W, self.H, self.variant = geom, binning, img, R, P, W, H, variant

    def _layout(self, fn, buf, n, k):
        off = (C.c_size_t * k)()
        getattr(lib(self.variant), fn)(C.c_void_p(buf.data_ptr()), C.c_size_t(n), off)
        return list(off)

    def parse(self):
        """Decode the carved state into numpy arrays."""
        P, R, N = self.P, self.R, self.W * self.H
        g = self.geom.cpu().numpy(); b = self.binning.cpu().numpy(); im = self.img.cpu().numpy()
        go = self._layout("ref_geom_layout", self.geom, P, 10)
        bo = self._layout("ref_binning_layout", self.binning, R, 5)
        io = self._layout("ref_image_layout", self.img, N, 3)

        def view(buf, off, dtype, count):
            return np.frombuffer(buf.tobytes()[off:off + np.dtype(dtype).itemsize * count], dtype=dtype).copy()
        tiles = ((self.W + 15) // 16) * ((self.H + 15) // 16)
        return dict(
            depths=view(g, go[0], np.float32, P), means2D=view(g, go[3], np.float32, 2 * P).reshape(P, 2),
            cov3D=view(g, go[4], np.float32, 6 * P).reshape(P, 6),
            conic_opacity=view(g, go[5], np.float32, 4 * P).reshape(P, 4),
            tiles_touched=view(g, go[7], np.uint32, P), point_offsets=view(g, go[9], np.uint32, P),
            point_list=view(b, bo[0], np.uint32, R), vals_unsorted=view(b, bo[1], np.uint32, R),
            keys_sorted=view(b, bo[2], np.uint64, R), keys_unsorted=view(b, bo[3], np.uint64, R),
            final_T=view(im, io[0], np.float32, N).reshape(self.H, self.W),
            n_contrib=view(im, io[1], np.uint32, N).reshape(self.H, self.W),
            ranges=view(im, io[2], np.uint32, 2 * tiles).reshape(tiles, 2), R=R)


def parse_light(st):
    """RefState.parse() without final_T / n_contrib and without copying the W*H-sized image buffer to the host: every field is
    sliced on the device first (the million-Gaussian test parses thousands of states)."""
    P, R = st.P, st.R
    go = st._layout("ref_geom_layout", st.geom, P, 10)
    bo = st._layout("ref_binning_layout", st.binning, R, 5)
    io = st._layout("ref_image_layout", st.img, st.W * st.H, 3)
    tiles = ((st.W + 15) // 16) * ((st.H + 15) // 16)

    def view(buf, off, dtype, count):
        n = np.dtype(dtype).itemsize * count
        return buf[off:off + n].cpu().numpy().view(dtype).copy()
    return dict(
        depths=view(st.geom, go[0], np.float32, P), means2D=view(st.geom, go[3], np.float32, 2 * P).reshape(P, 2),
        cov3D=view(st.geom, go[4], np.float32, 6 * P).reshape(P, 6), conic_opacity=view(st.geom, go[5], np.float32, 4 * P).reshape(P, 4),
        tiles_touched=view(st.geom, go[7], np.uint32, P), point_offsets=view(st.geom, go[9], np.uint32, P),
        point_list=view(st.binning, bo[0], np.uint32, R), vals_unsorted=view(st.binning, bo[1], np.uint32, R),
        keys_sorted=view(st.binning, bo[2], np.uint64, R), keys_unsorted=view(st.binning, bo[3], np.uint64, R),
        ranges=view(st.img, io[2], np.uint32, 2 * tiles).reshape(tiles, 2), R=R)


def rasterize_forward(variant, bg, means3D, colors_precomp, opacities, scales, rotations, scale_modifier,
                      cov3D_precomp, viewmatrix, projmatrix, tanfovx, tanfovy, H, W, sh, degree, campos,
                      prefiltered=False, antialiasing=False, debug=False, r_capacity=1 << 16):
    """RasterizeGaussiansCUDA (RAST/rasterize_points.cu:35-124)."""
    L = lib(variant)
    NC = VARIANT_CHANNELS[variant]
    if means3D.dim() != 2 or means3D.size(1) != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    P = means3D.size(0)
    dev = means3D.device
    out_color = torch.full((NC, H, W), 0.0, dtype=torch.float32, device=dev)
    out_invdepth = torch.full((1, H, W), 0.0, dtype=torch.float32, device=dev)
    radii = torch.full((P,), 0, dtype=torch.int32, device=dev)
    # +512: the reference carves at 128-B aligned absolute addresses
    geom = torch.empty(L.ref_required_geom(P) + 512, dtype=torch.uint8, device=dev)
    img = torch.empty(L.ref_required_image(W * H) + 512, dtype=torch.uint8, device=dev)
    binning = torch.empty(L.ref_required_binning(r_capacity) + 512, dtype=torch.uint8, device=dev)
    rendered = 0
    if P != 0:
        M = sh.size(1) if sh.numel() != 0 else 0
        used = (C.c_size_t * 3)()
        rendered = L.ref_forward(
            C.c_int(P), C.c_int(degree), C.c_int(M), _ptr(bg.contiguous()), C.c_int(W), C.c_int(H),
            _ptr(means3D.contiguous()), _ptr(sh.contiguous()), _ptr(colors_precomp.contiguous()),
            _ptr(opacities.contiguous()), _ptr(scales.contiguous()), C.c_float(scale_modifier),
            _ptr(rotations.contiguous()), _ptr(cov3D_precomp.contiguous()), _ptr(viewmatrix.contiguous()),
            _ptr(projmatrix.contiguous()), _ptr(campos.contiguous()), C.c_float(tanfovx), C.c_float(tanfovy),
            C.c_int(int(prefiltered)), _ptr(out_color), _ptr(out_invdepth), C.c_int(int(antialiasing)),
            _ptr(radii), C.c_int(int(debug)),
            _ptr(geom), C.c_size_t(geom.numel()), _ptr(binning), C.c_size_t(binning.numel()),
            _ptr(img), C.c_size_t(img.numel()), used)
        if rendered < 0:
            raise RuntimeError("reference forward failed (scratch arena too small?)")
    return rendered, out_color, radii, geom, binning, img, out_invdepth


def rasterize_backward(variant, bg, means3D, radii, colors_precomp, opacities, scales, rotations, scale_modifier,
                       cov3D_precomp, viewmatrix, projmatrix, tanfovx, tanfovy, dL_dout_color, dL_dout_invdepth,
                       sh, degree, campos, geom, R, binning, img, antialiasing=False, debug=False):
    """RasterizeGaussiansBackwardCUDA (RAST/rasterize_points.cu:126-223)."""
    L = lib(variant)
    NC = VARIANT_CHANNELS[variant]
    P = means3D.size(0)
    H, W = dL_dout_color.size(1), dL_dout_color.size(2)
    M = sh.size(1) if sh.numel() != 0 else 0
    o = dict(dtype=torch.float32, device=means3D.device)
    dL_dmeans3D = torch.zeros((P, 3), **o); dL_dmeans2D = torch.zeros((P, 3), **o)
    dL_dcolors = torch.zeros((P, NC), **o); dL_dconic = torch.zeros((P, 2, 2), **o)
    dL_dopacity = torch.zeros((P, 1), **o); dL_dcov3D = torch.zeros((P, 6), **o)
    dL_dsh = torch.zeros((P, M, NC), **o); dL_dscales = torch.zeros((P, 3), **o)
    dL_drotations = torch.zeros((P, 4), **o)
    dL_dinvdepths = torch.zeros((0, 1), **o)
    dinv_pix = None
    if dL_dout_invdepth is not None and dL_dout_invdepth.numel() != 0:
        dL_dinvdepths = torch.zeros((P, 1), **o)
        dinv_pix = dL_dout_invdepth.contiguous()
    if P != 0:
        rc = L.ref_backward(
            C.c_int(P), C.c_int(degree), C.c_int(M), C.c_int(R), _ptr(bg.contiguous()), C.c_int(W), C.c_int(H),
            _ptr(means3D.contiguous()), _ptr(sh.contiguous()), _ptr(colors_precomp.contiguous()),
            _ptr(opacities.contiguous()), _ptr(scales.contiguous()), C.c_float(scale_modifier),
            _ptr(rotations.contiguous()), _ptr(cov3D_precomp.contiguous()), _ptr(viewmatrix.contiguous()),
            _ptr(projmatrix.contiguous()), _ptr(campos.contiguous()), C.c_float(tanfovx), C.c_float(tanfovy),
            _ptr(radii), _ptr(geom), _ptr(binning), _ptr(img), _ptr(dL_dout_color.contiguous()), _ptr(dinv_pix),
            _ptr(dL_dmeans2D), _ptr(dL_dconic), _ptr(dL_dopacity), _ptr(dL_dcolors), _ptr(dL_dinvdepths),
            _ptr(dL_dmeans3D), _ptr(dL_dcov3D), _ptr(dL_dsh), _ptr(dL_dscales), _ptr(dL_drotations),
            C.c_int(int(antialiasing)), C.c_int(int(debug)))
        if rc != 0:
            raise RuntimeError("reference backward failed")
    return dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool
    antialiasing: bool


def make_function(variant):
    class _RasterizeGaussians(torch.autograd.Function):
        @staticmethod
        def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, rs):
            # the reference hands a 3-float bg to a kernel that reads C floats:
            # pad with zeros so the oracle is well defined.
            bg = torch.zeros(32, dtype=torch.float32, device=means3D.device)
            bg[:rs.bg.numel()] = rs.bg
            R, color, radii, geom, binning, img, invdepths = rasterize_forward(
                variant, bg, means3D, colors_precomp, opacities, scales, rotations, rs.scale_modifier,
                cov3Ds_precomp, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, rs.image_height,
                rs.image_width, sh, rs.sh_degree, rs.campos, rs.prefiltered, rs.antialiasing, rs.debug)
            ctx.rs, ctx.R, ctx.bg = rs, R, bg
            ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, opacities, geom, binning, img)
            ctx.mark_non_differentiable(radii)
            return color, radii, invdepths

        @staticmethod
        def backward(ctx, grad_out_color, _, grad_out_depth):
            rs = ctx.rs
            colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, opacities, geom, binning, img = ctx.saved_tensors
            g2d, gcol, gop, g3d, gcov, gsh, gsc, grot = rasterize_backward(
                variant, ctx.bg, means3D, radii, colors_precomp, opacities, scales, rotations, rs.scale_modifier,
                cov3Ds_precomp, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, grad_out_color, grad_out_depth,
                sh, rs.sh_degree, rs.campos, geom, ctx.R, binning, img, rs.antialiasing, rs.debug)
            return g3d, g2d, gsh, gcol, gop, gsc, grot, gcov, None
    return _RasterizeGaussians


_FUNCS = {}


class GaussianRasterizer(torch.nn.Module):
    def __init__(self, raster_settings, variant="h36m"):
        super().__init__()
        self.raster_settings = raster_settings
        self.variant = variant
        if variant not in _FUNCS:
            _FUNCS[variant] = make_function(variant)

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None, cov3D_precomp=None):
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        e = torch.Tensor([])
        shs = e if shs is None else shs
        colors_precomp = e if colors_precomp is None else colors_precomp
        scales = e if scales is None else scales
        rotations = e if rotations is None else rotations
        cov3D_precomp = e if cov3D_precomp is None else cov3D_precomp
        return _FUNCS[self.variant].apply(means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp, self.raster_settings)
