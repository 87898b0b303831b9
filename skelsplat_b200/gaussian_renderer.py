"""render_h36m / render_panoptic / render_op with the reference's signature and return dict
(gaussian_renderer/__init__.py:28-371), on top of the sm_100a rasteriser op.

Kept literally: the settings tuple, ``shs = pc.get_features`` ([J,1,C] one-hot features used directly
as per-Gaussian channels), ``clamp(0, 1)``, and the returned keys.  One deliberate difference:
``visibility_filter`` is computed lazily (``(radii > 0).nonzero()`` forces a host sync per render in
the reference, :133, and nothing in train.py consumes it); access it as ``out["visibility_filter"]``
and it is materialised on demand.
"""
import math

import torch

import diff_gaussian_rasterization_h36m as _h36m
import diff_gaussian_rasterization_op as _op
import diff_gaussian_rasterization_panoptic as _pan


class _LazyDict(dict):
    def __getitem__(self, k):
        v = dict.__getitem__(self, k)
        if callable(v) and k == "visibility_filter":
            v = v()
            dict.__setitem__(self, k, v)
        return v


def _render(mod, viewpoint_camera, pc, pipe, bg_color, scaling_modifier=1.0, separate_sh=False, override_color=None, use_trained_exp=False):
    screenspace_points = torch.zeros_like(pc.get_xyz, dtype=pc.get_xyz.dtype, requires_grad=True) + 0
    try:
        screenspace_points.retain_grad()
    except Exception:
        pass
    tanfovx = math.tan(viewpoint_camera.FoVx * 0.5)
    tanfovy = math.tan(viewpoint_camera.FoVy * 0.5)
    raster_settings = mod.GaussianRasterizationSettings(
        image_height=int(viewpoint_camera.image_height), image_width=int(viewpoint_camera.image_width),
        tanfovx=tanfovx, tanfovy=tanfovy, bg=bg_color, scale_modifier=scaling_modifier,
        viewmatrix=viewpoint_camera.world_view_transform, projmatrix=viewpoint_camera.full_proj_transform,
        sh_degree=pc.active_sh_degree, campos=viewpoint_camera.camera_center, prefiltered=False,
        debug=getattr(pipe, "debug", False), antialiasing=getattr(pipe, "antialiasing", False))
    rasterizer = mod.GaussianRasterizer(raster_settings=raster_settings)
    scales = rotations = cov3D_precomp = None
    if getattr(pipe, "compute_cov3D_python", False):
        cov3D_precomp = pc.get_covariance(scaling_modifier)
    else:
        scales, rotations = pc.get_scaling, pc.get_rotation
    shs = colors_precomp = None
    if override_color is None:
        if getattr(pipe, "convert_SHs_python", False):
            raise NotImplementedError("convert_SHs_python is false in every SkelSplat config (SH evaluation is out of scope)")
        shs = pc.get_features
    else:
        colors_precomp = override_color
    rendered_image, radii, depth_image = rasterizer(
        means3D=pc.get_xyz, means2D=screenspace_points, shs=shs, colors_precomp=colors_precomp,
        opacities=pc.get_opacity, scales=scales, rotations=rotations, cov3D_precomp=cov3D_precomp)
    rendered_image = rendered_image.clamp(0, 1)
    return _LazyDict({"render": rendered_image, "viewspace_points": screenspace_points,
                      "visibility_filter": (lambda: (radii > 0).nonzero()), "radii": radii, "depth": depth_image})


def render_h36m(viewpoint_camera, pc, pipe, bg_color, scaling_modifier=1.0, separate_sh=False, override_color=None, use_trained_exp=False):
    return _render(_h36m, viewpoint_camera, pc, pipe, bg_color, scaling_modifier, separate_sh, override_color, use_trained_exp)


def render_panoptic(viewpoint_camera, pc, pipe, bg_color, scaling_modifier=1.0, separate_sh=False, override_color=None, use_trained_exp=False):
    return _render(_pan, viewpoint_camera, pc, pipe, bg_color, scaling_modifier, separate_sh, override_color, use_trained_exp)


def render_op(viewpoint_camera, pc, pipe, bg_color, scaling_modifier=1.0, separate_sh=False, override_color=None, use_trained_exp=False):
    return _render(_op, viewpoint_camera, pc, pipe, bg_color, scaling_modifier, separate_sh, override_color, use_trained_exp)


render_functions = {
    "diff-gaussian-rasterization-h36m": render_h36m,
    "diff-gaussian-rasterization-panoptic": render_panoptic,
    "diff-gaussian-rasterization-op": render_op,
}
