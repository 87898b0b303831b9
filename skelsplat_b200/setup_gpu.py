"""Per-frame setup on the GPU: batched DLT initial guess and GT heatmap ROIs.

Host counterparts (numpy, used by the CPU tests and as the readable specification): ``triangulation.triangulate_poses``
and ``heatmaps.generate_heatmap_rois``.  With these two kernels a whole sequence goes detections -> initial guess ->
heatmap ROIs -> fused optimisation without any per-frame host work (``pack_sequence_gpu``).
"""
import ctypes as C

import numpy as np
import torch

from . import lib as _L
from .cameras import cameras_extent
from .configs import SceneConfig
from .trainer import PackedSequence, camera_tensors, initial_raw_state


def triangulate_dlt(P_list, poses_2d, device="cuda"):
    """[F,J,3] float64 tensor from V projection matrices K[R|t] and detections [F,V,J,2] (triangulation.py:122-150)."""
    L = _L.lib()
    P = (P_list if torch.is_tensor(P_list) else torch.as_tensor(np.asarray(P_list, np.float64))).to(device=device, dtype=torch.float64).contiguous()
    d = (poses_2d if torch.is_tensor(poses_2d) else torch.as_tensor(np.asarray(poses_2d))).to(device=device, dtype=torch.float64).contiguous()
    F, V, J = d.shape[0], d.shape[1], d.shape[2]
    out = torch.empty((F, J, 3), dtype=torch.float64, device=device)
    _L.check(L.ssb_triangulate_dlt(C.c_int(F), C.c_int(V), C.c_int(J), _L.ptr(P), _L.ptr(d), _L.ptr(out), _L.current_stream()),
             "ssb_triangulate_dlt")
    return out


def _roi_rects(cfg, cams_struct, xyz, scaling, rotation, p2d, rect, sigma, center, size):
    L = _L.lib()
    F, J = xyz.shape[0], cfg.n_joints
    _L.check(L.ssb_heatmap_roi_rects(C.c_int(F), C.c_int(J), C.byref(cams_struct), _L.ptr(xyz), _L.ptr(scaling), _L.ptr(rotation), _L.ptr(p2d),
                                     C.c_float(1.0), _L.ptr(rect), _L.ptr(sigma), _L.ptr(center), _L.ptr(size), _L.current_stream()),
             "ssb_heatmap_roi_rects")


def generate_heatmap_rois_gpu(cfg: SceneConfig, vm, pm, dims, tanfov, Wmax, Hmax, xyz, scaling, rotation, poses_2d):
    """Device tensors in, device tensors out: (roi_rect [F,V,J,4] int32, roi_offset [F,V,J] int64, roi_data [total] fp32).
    Exactly sized output: one host read of the packed size (use ``generate_heatmap_rois_into`` for the sync-free form)."""
    L = _L.lib()
    F, J, V = xyz.shape[0], cfg.n_joints, cfg.nviews
    dev = xyz.device
    cams = _L.Cameras(V, _L.ptr(vm), _L.ptr(pm), _L.ptr(dims), _L.ptr(tanfov), Wmax, Hmax, 0.0, 0.0, 0)
    rect = torch.empty((F, V, J, 4), dtype=torch.int32, device=dev)
    sigma = torch.empty((F, V, J, 2), dtype=torch.float32, device=dev)
    center = torch.empty((F, V, J, 2), dtype=torch.int32, device=dev)
    size = torch.empty((F, V, J), dtype=torch.int64, device=dev)
    offset = torch.empty((F, V, J), dtype=torch.int64, device=dev)
    total = torch.zeros(1, dtype=torch.int64, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    p2d = poses_2d.to(torch.float32).contiguous()
    _roi_rects(cfg, cams, xyz, scaling, rotation, p2d, rect, sigma, center, size)
    _L.check(L.ssb_heatmap_roi_offsets(C.c_int64(F * V * J), _L.ptr(size), _L.ptr(offset), _L.ptr(total), _L.current_stream()),
             "ssb_heatmap_roi_offsets")
    data = torch.empty(int(total.item()), dtype=torch.float32, device=dev)      # the one host sync: sizes the packed buffer
    _L.check(L.ssb_heatmap_roi_fill(C.c_int(F), C.c_int(J), C.byref(cams), _L.ptr(rect), _L.ptr(sigma), _L.ptr(center), _L.ptr(offset),
                                    _L.ptr(data), C.c_int64(-1), _L.ptr(status), _L.current_stream()), "ssb_heatmap_roi_fill")
    if int(status.item()) & 4:
        raise _L.SkelSplatLibraryError("heatmap patch wider than 256 px (sigma > 31 px): outside the supported regime")
    return rect, offset, data


def generate_heatmap_rois_into(cfg: SceneConfig, vm, pm, dims, tanfov, Wmax, Hmax, xyz, scaling, rotation, p2d,
                               rect, sigma, center, size, offset, data, status):
    """Sync-free form for pipelines: every buffer is caller-provided; ``data`` is a capacity buffer and ``status`` (int32[1],
    zeroed by the caller) receives SSB_STATUS_ROI_OVERFLOW / SSB_STATUS_ROI_TOO_WIDE if a patch could not be written."""
    L = _L.lib()
    F, J, V = xyz.shape[0], cfg.n_joints, cfg.nviews
    cams = _L.Cameras(V, _L.ptr(vm), _L.ptr(pm), _L.ptr(dims), _L.ptr(tanfov), Wmax, Hmax, 0.0, 0.0, 0)
    _roi_rects(cfg, cams, xyz, scaling, rotation, p2d, rect, sigma, center, size)
    _L.check(L.ssb_heatmap_roi_offsets(C.c_int64(F * V * J), _L.ptr(size), _L.ptr(offset), None, _L.current_stream()),
             "ssb_heatmap_roi_offsets")
    _L.check(L.ssb_heatmap_roi_fill(C.c_int(F), C.c_int(J), C.byref(cams), _L.ptr(rect), _L.ptr(sigma), _L.ptr(center), _L.ptr(offset),
                                    _L.ptr(data), C.c_int64(data.numel()), _L.ptr(status), _L.current_stream()), "ssb_heatmap_roi_fill")


def pack_sequence_gpu(cfg: SceneConfig, cams, poses_2d, poses_init=None, device="cuda") -> PackedSequence:
    """detections [F,V,J,2] (+ optional initial poses) -> PackedSequence, entirely on the GPU.
    Without ``poses_init`` the initial guess is the DLT triangulation of the detections (BASELINE config 1)."""
    d2 = (poses_2d if torch.is_tensor(poses_2d) else torch.as_tensor(np.asarray(poses_2d))).to(device)
    F = d2.shape[0]
    if poses_init is None:
        init = triangulate_dlt([c.P3x4() for c in cams], d2, device).to(torch.float32)
    elif torch.is_tensor(poses_init):
        init = poses_init.to(device=device, dtype=torch.float32)
    else:
        init = torch.as_tensor(np.asarray(poses_init, np.float32)).to(device)
    _, scal, rot, opa = initial_raw_state(cfg, np.zeros((F, cfg.n_joints, 3), np.float32))
    scaling, rotation, opacity = (torch.from_numpy(a).to(device) for a in (scal, rot, opa))
    vm, pm, dims, tanfov = camera_tensors(cams, device)
    Wmax, Hmax = max(c.image_width for c in cams), max(c.image_height for c in cams)
    xyz = init.contiguous()
    rect, offset, data = generate_heatmap_rois_gpu(cfg, vm, pm, dims, tanfov, Wmax, Hmax, xyz, scaling, rotation, d2)
    return PackedSequence(cfg=cfg, n_frames=F, xyz=xyz.clone(), scaling=scaling, rotation=rotation, opacity=opacity, viewmatrix=vm,
                          projmatrix=pm, dims=dims, tanfov=tanfov, roi_rect=rect, roi_offset=offset, roi_data=data,
                          spatial_lr_scale=cameras_extent(cams), Wmax=Wmax, Hmax=Hmax)
