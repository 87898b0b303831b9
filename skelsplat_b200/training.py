"""train.py's per-frame loop on the DROP-IN surface (render_* + losses + GaussianModel + torch Adam).

This is the reference's own iteration body (train.py:130-233) with the reference's Python modules
replaced one for one by skelsplat_b200's mirrors: every dense tensor of the reference contract is
still materialised, so it measures what a user gets by swapping the packages without touching
train.py.  The fused path (trainer.optimize_sequence) is the B200-native form of the same loop.
Dropped: tensorboard / logging / ``empty_cache`` / the per-4-iteration synchronises (train.py:184-213,
260-276) and ``loss.item()`` (155-158; NotStopping ignores it) -- none of them changes a result.
"""
from types import SimpleNamespace

import numpy as np
import torch

from .cameras import cameras_extent
from .gaussian_model import GaussianModel
from .gaussian_renderer import render_functions
from .heatmaps import generate_heatmap_rois, rois_to_dense
from .loss_utils import consistency_losses, losses


class TorchCamera:
    """Device-tensor view of a cameras.ViewCamera with the attribute names render_* reads (scene/cameras.py)."""

    def __init__(self, cam, device="cuda"):
        self.uid = cam.uid
        self.image_width, self.image_height = cam.image_width, cam.image_height
        self.FoVx, self.FoVy = cam.FoVx, cam.FoVy
        self.world_view_transform = torch.from_numpy(cam.world_view_transform).to(device)
        self.full_proj_transform = torch.from_numpy(cam.full_proj_transform).to(device)
        self.camera_center = torch.from_numpy(cam.camera_center).to(device)


def optimise_frame_dropin(frame, cams, cfg, heatmaps_dense=None, device="cuda", iterations=None, modules=None, init_state=None,
                          return_state=False, spatial_lr_scale=None):
    """One frame through the drop-in API; returns final xyz [J,3] float32 numpy.

    ``frame`` needs ``pose_3d_init`` and ``poses_2d`` (the latter only to build the GT heatmaps when ``heatmaps_dense`` is None).
    ``modules``: optional (GaussianModel, render_functions, losses, consistency_losses) to run the same loop body on another
    implementation of the same surface -- the tests pass the REFERENCE's own classes and functions here (tests/ref_import.py),
    which is how "the reference's Python runs unchanged on the drop-in packages" is checked."""
    iterations = cfg.iterations if iterations is None else iterations
    extent = cameras_extent(cams) if spatial_lr_scale is None else spatial_lr_scale
    GM, rfuncs, loss_table, cons_table = modules if modules is not None else (GaussianModel, render_functions, losses, consistency_losses)
    ref_surface = modules is not None
    opt = SimpleNamespace(position_lr_init=cfg.position_lr_init, position_lr_final=cfg.position_lr_final,
                          position_lr_delay_mult=cfg.position_lr_delay_mult, position_lr_max_steps=cfg.position_lr_max_steps,
                          feature_lr=cfg.feature_lr, opacity_lr=cfg.opacity_lr, scaling_lr=cfg.scaling_lr,
                          rotation_lr=cfg.rotation_lr, percent_dense=0.01)
    pipe = SimpleNamespace(debug=False, antialiasing=cfg.antialiasing, compute_cov3D_python=False, convert_SHs_python=False)
    data_root = "data/" + cfg.name
    if ref_surface:      # the reference's signatures: GaussianModel(sh_degree, optimizer_type); create_from_pcd(pcd: BasicPointCloud, cam_infos, ...)
        gaussians = GM(1, "default")
        pts = np.asarray(frame.pose_3d_init, np.float32)
        pcd = SimpleNamespace(points=pts, colors=np.zeros_like(pts), normals=np.zeros_like(pts))
        cam_infos = [SimpleNamespace(image_name=f"cam{c.uid}") for c in cams]
        opt.exposure_lr_init, opt.exposure_lr_final, opt.exposure_lr_delay_steps, opt.exposure_lr_delay_mult, opt.iterations = 0.01, 0.001, 0, 0.0, iterations
        gaussians.create_from_pcd(pcd, cam_infos, extent, cfg.opacity_on, cfg.scaling, cfg.n_joints, cfg.scaling_modifier, cfg.name)
    else:
        gaussians = GM(1, "default", device)
        gaussians.create_from_pcd(np.asarray(frame.pose_3d_init, np.float32), cams, extent, cfg.opacity_on, cfg.scaling,
                                  cfg.n_joints, cfg.scaling_modifier, cfg.name)
    if init_state is not None:      # (scaling_raw [J,3], rotation_raw [J,4]): start from a given raw state (dense fallback of the fused path)
        with torch.no_grad():
            gaussians._scaling.copy_(torch.as_tensor(init_state[0]).to(device)); gaussians._rotation.copy_(torch.as_tensor(init_state[1]).to(device))
    gaussians.training_setup(opt)
    tcams = [TorchCamera(c, device) for c in cams]
    if heatmaps_dense is None:
        rois = generate_heatmap_rois(np.asarray(frame.pose_3d_init), frame.poses_2d, cams,
                                     gaussians._scaling.detach().cpu().numpy(), gaussians._rotation.detach().cpu().numpy())
        heatmaps_dense = [torch.from_numpy(rois_to_dense(rois, v)).to(device) for v in range(len(cams))]
    render = rfuncs[cfg.rendering]
    opt_criterion = loss_table[cfg.loss_function]
    consistency_criterion = cons_table[cfg.consistency_loss]
    bg = torch.tensor([0, 0, 0], dtype=torch.float32, device=device)
    accumulated_grads = torch.zeros((len(tcams),) + tuple(gaussians.get_xyz.shape), device=device)
    # gt_2d argument of the loss table's signature (utils/loss_utils.py:67,86): only the never-configured soft-argmax losses read it
    poses_2d = torch.as_tensor(np.asarray(frame.poses_2d)) if frame.poses_2d is not None else torch.zeros((len(cams), cfg.n_joints, 2))
    for iteration in range(1, iterations + 1):
        gaussians.update_learning_rate(iteration)
        idx = (iteration - 1) % len(tcams)
        render_pkg = render(tcams[idx], gaussians, pipe, bg)
        image = render_pkg["render"]
        l2_loss, error = opt_criterion(image, heatmaps_dense[idx], poses_2d[idx, :, :2], cfg.lambda_loss_function, reduction="mean")
        loss = l2_loss + consistency_criterion(gaussians.get_xyz, data_root, reduction="mean") * cfg.lambda_consistency
        params = [gaussians.get_xyz, gaussians._scaling, gaussians._rotation, gaussians._opacity]
        grads = torch.autograd.grad(loss, params)
        accumulated_grads[idx, ...] = grads[0]
        gaussians._scaling.grad, gaussians._rotation.grad, gaussians._opacity.grad = grads[1], grads[2], grads[3]
        if iteration % cfg.accumulation_steps == 0:
            gaussians.get_xyz.grad = accumulated_grads.to(gaussians.get_xyz.dtype).mean(dim=0)
            with torch.no_grad():
                gaussians.optimizer.step()
                gaussians.optimizer.zero_grad(set_to_none=True)
    if return_state:
        return tuple(t.detach().clone() for t in (gaussians._xyz, gaussians._scaling, gaussians._rotation, gaussians._opacity))
    return gaussians._xyz.detach().cpu().numpy().copy()
