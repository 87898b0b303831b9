"""Host-side mirror of the reference's rasteriser op, backed by the sm_100a C-ABI library.

Same names, argument order and error behaviour as the reference's
``diff_gaussian_rasterization_{h36m,panoptic,op}`` packages
(RAST/diff_gaussian_rasterization_h36m/__init__.py:44-207):
``GaussianRasterizationSettings`` (13-field NamedTuple), ``GaussianRasterizer`` with
``forward(means3D, means2D, opacities, shs, colors_precomp, scales, rotations, cov3D_precomp)``
-> ``(color[C,H,W], radii[P] int32, invdepth[1,H,W])`` and ``markVisible(positions)``.

Differences, all deliberate:
  * no host synchronisation: ``num_rendered`` stays on the device (read it with
    ``RasterState.num_rendered`` when a test needs it);
  * the gradient w.r.t. the per-Gaussian features is returned in the slot of the tensor
    that supplied them (``shs`` in SkelSplat); the reference returns garbage for ``shs``
    (uninitialised ``clamped`` in its SH backward);
  * backward is deterministic (no atomics).
Batched entry points (`rasterize_batched`) expose the frames x views form used by the bench.
"""
import ctypes as C
import os
from typing import NamedTuple, Optional

import numpy as np
import torch
import torch.nn as nn

from . import lib as _L

DEFAULT_R_CAPACITY = int(os.environ.get("SKELSPLAT_B200_R_CAPACITY", 2048))     # (Gaussian,tile) pairs per view of the drop-in op
MAX_R_CAPACITY = 1 << 14                                                        # ssb_rasterize_forward's upper bound


class _OverflowWatch:
    """Capacity overflow of the drop-in op, made loud WITHOUT a per-render host sync: the forward kernels NaN-fill the image of
    an overflowed view (so every loss over it is NaN), and the 32-byte state header is copied to pinned host memory behind
    the kernels; backward -- by which time the copy has long completed -- reads it and raises."""
    N = 64

    def __init__(self):
        self.slots = None
        self.i = 0
        self.gen = [0] * self.N          # a slot is reused after N renders: a ticket older than that is not checked (its data is gone)

    def post(self, state_buf):
        if torch.cuda.is_current_stream_capturing():
            return None
        if self.slots is None:
            self.slots = [(torch.zeros(8, dtype=torch.int32).pin_memory(), torch.cuda.Event()) for _ in range(self.N)]
        self.i = (self.i + 1) % self.N
        host, ev = self.slots[self.i]
        ev.synchronize()                 # the slot's previous copy (N renders ago) has long completed; never overwrite one in flight
        self.gen[self.i] += 1
        host.copy_(state_buf[:32].view(torch.int32), non_blocking=True)
        ev.record()
        return (self.i, host, ev, self.gen[self.i])

    def check(self, ticket, rcap):
        if ticket is None:
            return
        i, host, ev, gen = ticket
        if self.gen[i] != gen:           # more than N renders between this forward and its backward: the NaN-filled image is the only signal left
            return
        ev.synchronize()
        if int(host[2]) != 0:
            raise _L.SkelSplatLibraryError(
                f"rasteriser capacity exceeded: this view needs {int(host[7])} (Gaussian,tile) pairs, r_capacity is {rcap}; the rendered image was "
                f"NaN-filled.  Raise skelsplat_b200.rasterizer.DEFAULT_R_CAPACITY (or SKELSPLAT_B200_R_CAPACITY) up to {MAX_R_CAPACITY}")


_overflow_watch = _OverflowWatch()


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


class RasterState:
    """Opaque per-view state (the reference's geomBuffer/binningBuffer/imgBuffer + R) with debug
    accessors for the bit-exact stage tests."""

    def __init__(self, buf, B, P, W, H, rcap):
        self.buf, self.B, self.P, self.W, self.H, self.rcap = buf, B, P, W, H, rcap
        self.stride = _L.state_bytes(P, W, H, rcap)

    def _field(self, b, f, dtype, count):
        off = b * self.stride + _L.state_field_offset(self.P, self.W, self.H, self.rcap, f)
        raw = self.buf[off:off + count * np.dtype(dtype).itemsize].cpu().numpy()
        return raw.view(dtype).copy()

    def header(self, b=0):
        return self._field(b, _L.F_HEADER, np.int32, 8)

    def num_rendered(self, b=0):
        return int(self.header(b)[0])

    def check(self):
        """Raise if any view overflowed r_capacity (device-side status word)."""
        for b in range(self.B):
            h = self.header(b)
            if h[2] != 0:
                raise _L.SkelSplatLibraryError(f"view {b}: {h[7]} (Gaussian,tile) pairs exceed r_capacity={self.rcap}")

    def parse(self, b=0, W=None, H=None):
        P, rc = self.P, self.rcap
        W = self.W if W is None else W
        H = self.H if H is None else H
        h = self.header(b)
        R, nact = int(h[0]), int(h[1])
        tiles = ((W + 15) // 16) * ((H + 15) // 16)
        f = self._field
        return dict(
            R=R, n_active=nact, status=int(h[2]),
            depths=f(b, _L.F_DEPTHS, np.float32, P), means2D=f(b, _L.F_MEANS2D, np.float32, 2 * P).reshape(P, 2),
            conic_opacity=f(b, _L.F_CONIC_OPACITY, np.float32, 4 * P).reshape(P, 4),
            cov3D=f(b, _L.F_COV3D, np.float32, 6 * P).reshape(P, 6),
            tiles_touched=f(b, _L.F_TILES_TOUCHED, np.uint32, P), point_offsets=f(b, _L.F_POINT_OFFSETS, np.uint32, P),
            rects=f(b, _L.F_RECTS, np.uint32, 4 * P).reshape(P, 4),
            keys_unsorted=f(b, _L.F_KEYS_UNSORTED, np.uint64, rc)[:R], vals_unsorted=f(b, _L.F_VALS_UNSORTED, np.uint32, rc)[:R],
            keys_sorted=f(b, _L.F_KEYS_SORTED, np.uint64, rc)[:R], point_list=f(b, _L.F_POINT_LIST, np.uint32, rc)[:R],
            inv_pos=f(b, _L.F_INV_POS, np.uint32, rc)[:R], tile_ids=f(b, _L.F_TILE_IDS, np.uint32, rc)[:nact],
            tile_ranges=f(b, _L.F_TILE_RANGES, np.uint32, 2 * rc).reshape(rc, 2)[:nact],
            ranges=f(b, _L.F_RANGES, np.uint32, 2 * tiles).reshape(tiles, 2))


def _f32c(t):
    return t.contiguous() if t.dtype == torch.float32 else t.float().contiguous()


def _make_structs(P, Cch, means3D, scales, rotations, cov3D, opacities, features, features_per_frame, scale_modifier,
                  n_views, viewmatrix, projmatrix, dims, tanfov, W0, H0, tfx0, tfy0, antialiasing):
    g = _L.Gaussians(P, Cch, _L.ptr(means3D), _L.ptr(scales), _L.ptr(rotations), _L.ptr(cov3D), _L.ptr(opacities),
                     _L.ptr(features), int(features_per_frame), float(scale_modifier))
    c = _L.Cameras(n_views, _L.ptr(viewmatrix), _L.ptr(projmatrix), _L.ptr(dims), _L.ptr(tanfov), int(W0), int(H0),
                   float(tfx0), float(tfy0), int(bool(antialiasing)))
    return g, c


def rasterize_batched(means3D, scales, rotations, opacities, features, viewmatrix, projmatrix, W, H, tanfovx, tanfovy,
                      scale_modifier=1.0, cov3D_precomp=None, antialiasing=False, r_capacity=DEFAULT_R_CAPACITY,
                      dims=None, tanfov=None, color_offsets=None, invdepth_offsets=None, out_color=None,
                      out_invdepth=None, render_invdepth=True, state=None):
    """Frames x views forward.  means3D [F,P,3], scales/rotations [F,P,3|4] (activated), opacities [F,P],
    features [F,P,C] or [P,C]; viewmatrix/projmatrix [V,4,4]; uniform W x H unless dims/tanfov [V,2] and the
    offsets are given (then W,H must be the maxima).  Returns (color [B,C,H,W], radii [B,P], invdepth [B,1,H,W], RasterState)."""
    L = _L.lib()
    F, P = means3D.shape[0], means3D.shape[1]
    V = viewmatrix.shape[0]
    B = F * V
    Cch = features.shape[-1]
    dev = means3D.device
    if out_color is None:
        out_color = torch.empty((B, Cch, H, W), dtype=torch.float32, device=dev)
    if out_invdepth is None and render_invdepth:
        out_invdepth = torch.empty((B, 1, H, W), dtype=torch.float32, device=dev)
    radii = torch.empty((B, P), dtype=torch.int32, device=dev)
    if state is None:
        state = torch.empty(B * _L.state_bytes(P, W, H, r_capacity), dtype=torch.uint8, device=dev)
    g, c = _make_structs(P, Cch, means3D, scales, rotations, cov3D_precomp, opacities, features, features.dim() == 3,
                         scale_modifier, V, viewmatrix, projmatrix, dims, tanfov, W, H, tanfovx, tanfovy, antialiasing)
    rc = L.ssb_rasterize_forward(C.c_int(F), C.byref(g), C.byref(c), C.c_int(r_capacity), _L.ptr(out_color),
                                 _L.ptr(color_offsets), _L.ptr(out_invdepth), _L.ptr(invdepth_offsets), _L.ptr(radii),
                                 _L.ptr(state), _L.current_stream())
    _L.check(rc, "ssb_rasterize_forward")
    return out_color, radii, out_invdepth, RasterState(state, B, P, W, H, r_capacity)


def rasterize_batched_backward(st: RasterState, means3D, scales, rotations, opacities, features, viewmatrix, projmatrix,
                               W, H, tanfovx, tanfovy, dL_dcolor, dL_dinvdepth=None, scale_modifier=1.0,
                               cov3D_precomp=None, antialiasing=False, dims=None, tanfov=None, color_offsets=None,
                               invdepth_offsets=None, want=("means3D", "means2D", "scales", "rotations", "opacity", "features", "cov3D", "conic")):
    """Frames x views backward; returns a dict of per-view gradients [B,P,...]."""
    L = _L.lib()
    F, P = means3D.shape[0], means3D.shape[1]
    V = viewmatrix.shape[0]
    B = F * V
    Cch = features.shape[-1]
    dev = means3D.device
    shapes = dict(means3D=(B, P, 3), means2D=(B, P, 3), scales=(B, P, 3), rotations=(B, P, 4), opacity=(B, P, 1),
                  features=(B, P, Cch), cov3D=(B, P, 6), conic=(B, P, 2, 2))
    out = {k: (torch.empty(shapes[k], dtype=torch.float32, device=dev) if k in want else None) for k in shapes}
    scratch = torch.empty(B * _L.backward_scratch_bytes(Cch, st.rcap), dtype=torch.uint8, device=dev)
    g, c = _make_structs(P, Cch, means3D, scales, rotations, cov3D_precomp, opacities, features, features.dim() == 3,
                         scale_modifier, V, viewmatrix, projmatrix, dims, tanfov, W, H, tanfovx, tanfovy, antialiasing)
    rc = L.ssb_rasterize_backward(C.c_int(F), C.byref(g), C.byref(c), C.c_int(st.rcap), _L.ptr(dL_dcolor),
                                  _L.ptr(color_offsets), _L.ptr(dL_dinvdepth), _L.ptr(invdepth_offsets), _L.ptr(st.buf),
                                  _L.ptr(scratch), _L.ptr(out["means3D"]), _L.ptr(out["means2D"]), _L.ptr(out["scales"]),
                                  _L.ptr(out["rotations"]), _L.ptr(out["opacity"]), _L.ptr(out["features"]),
                                  _L.ptr(out["cov3D"]), _L.ptr(out["conic"]), _L.current_stream())
    _L.check(rc, "ssb_rasterize_backward")
    return out


class _RasterizeGaussians(torch.autograd.Function):
    """Single-view autograd op with the reference's signature (RAST/.../__init__.py:44-141)."""

    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, raster_settings, n_channels):
        rs = raster_settings
        ctx.set_materialize_grads(False)
        if means3D.dim() != 2 or means3D.size(1) != 3:
            raise RuntimeError("means3D must have dimensions (num_points, 3)")   # RAST/rasterize_points.cu:58-60
        P = means3D.size(0)
        H, W = int(rs.image_height), int(rs.image_width)
        dev = means3D.device
        if P == 0:   # RAST/rasterize_points.cu:88: nothing is launched, outputs are zeros
            ctx.empty = True
            return (torch.zeros((n_channels, H, W), dtype=torch.float32, device=dev), torch.zeros((0,), dtype=torch.int32, device=dev),
                    torch.zeros((1, H, W), dtype=torch.float32, device=dev))
        use_sh = sh.numel() != 0
        feats = _f32c(sh if use_sh else colors_precomp).reshape(P, -1)
        if feats.shape[1] != n_channels:
            raise RuntimeError(f"this rasteriser variant renders {n_channels} channels, got features with {feats.shape[1]}")
        has_cov = cov3Ds_precomp.numel() != 0
        m3 = _f32c(means3D).unsqueeze(0)
        sc = None if has_cov else _f32c(scales).unsqueeze(0)
        ro = None if has_cov else _f32c(rotations).unsqueeze(0)
        cv = _f32c(cov3Ds_precomp).unsqueeze(0) if has_cov else None
        op = _f32c(opacities).reshape(1, P)
        vm = _f32c(rs.viewmatrix).reshape(1, 4, 4)
        pm = _f32c(rs.projmatrix).reshape(1, 4, 4)
        color, radii, invd, st = rasterize_batched(m3, sc, ro, op, feats, vm, pm, W, H, rs.tanfovx, rs.tanfovy,
                                                   rs.scale_modifier, cv, rs.antialiasing, r_capacity=DEFAULT_R_CAPACITY)
        ctx.overflow_ticket = _overflow_watch.post(st.buf)
        if rs.debug:
            torch.cuda.synchronize()
            st.check()
        ctx.empty = False
        ctx.rs, ctx.st, ctx.use_sh, ctx.has_cov = rs, st, use_sh, has_cov
        ctx.shapes = (sh.shape, colors_precomp.shape, opacities.shape)
        ctx.save_for_backward(m3, sc, ro, cv, op, feats, vm, pm)
        radii = radii.reshape(P)
        ctx.mark_non_differentiable(radii)
        return color.reshape(n_channels, H, W), radii, invd.reshape(1, H, W)

    @staticmethod
    def backward(ctx, grad_out_color, _, grad_out_depth):
        if ctx.empty or grad_out_color is None:
            return (None,) * 10
        rs, st = ctx.rs, ctx.st
        _overflow_watch.check(ctx.overflow_ticket, st.rcap)
        m3, sc, ro, cv, op, feats, vm, pm = ctx.saved_tensors
        H, W = int(rs.image_height), int(rs.image_width)
        g = rasterize_batched_backward(st, m3, sc, ro, op, feats, vm, pm, W, H, rs.tanfovx, rs.tanfovy,
                                       _f32c(grad_out_color), None if grad_out_depth is None else _f32c(grad_out_depth),
                                       rs.scale_modifier, cv, rs.antialiasing,
                                       want=("means3D", "means2D", "scales", "rotations", "opacity", "features", "cov3D"))
        sh_shape, col_shape, op_shape = ctx.shapes
        gfeat = g["features"][0]
        grad_sh = gfeat.reshape(sh_shape) if ctx.use_sh else None
        grad_col = None if ctx.use_sh else gfeat.reshape(col_shape)
        grad_scales = None if ctx.has_cov else g["scales"][0]
        grad_rot = None if ctx.has_cov else g["rotations"][0]
        grad_cov = g["cov3D"][0] if ctx.has_cov else None
        return (g["means3D"][0], g["means2D"][0], grad_sh, grad_col, g["opacity"][0].reshape(op_shape),
                grad_scales, grad_rot, grad_cov, None, None)


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, raster_settings, n_channels):
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                                     raster_settings, n_channels)


class GaussianRasterizer(nn.Module):
    NUM_CHANNELS: Optional[int] = None     # set by the per-variant packages (17 / 19 / 15)

    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        with torch.no_grad():
            P = positions.shape[0]
            present = torch.zeros((P,), dtype=torch.bool, device=positions.device)
            if P:
                rc = _L.lib().ssb_mark_visible(C.c_int(P), _L.ptr(_f32c(positions)), _L.ptr(_f32c(self.raster_settings.viewmatrix)),
                                               _L.ptr(_f32c(self.raster_settings.projmatrix)), _L.ptr(present), _L.current_stream())
                _L.check(rc, "ssb_mark_visible")
        return present

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None, cov3D_precomp=None):
        raster_settings = self.raster_settings
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        empty = torch.Tensor([])
        shs = empty if shs is None else shs
        colors_precomp = empty if colors_precomp is None else colors_precomp
        scales = empty if scales is None else scales
        rotations = empty if rotations is None else rotations
        cov3D_precomp = empty if cov3D_precomp is None else cov3D_precomp
        n_channels = self.NUM_CHANNELS
        if n_channels is None:
            n_channels = (shs if shs.numel() else colors_precomp).reshape(means3D.shape[0], -1).shape[1]
        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp,
                                   raster_settings, n_channels)
