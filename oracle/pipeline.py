"""TEST INFRASTRUCTURE ONLY -- restatement of the reference's per-frame optimisation loop.

The reference's Python modules cannot be imported here (hydra, omegaconf, cupy,
tensordict, plyfile, open3d are absent), so this file restates the
loop literally, in plain torch, from:
  * train.py:56-233                      frame loop / iteration loop / grad bookkeeping / step
  * scene/gaussian_model.py:32-47,102-143,149-248   state, activations, Adam groups, LR schedule
  * utils/general_utils.py:27-28,38-71,175-304      inverse_sigmoid, LR func, generate_heatmaps
  * utils/loss_utils.py:67-127,226-250,257-300      losses, limb consistency, conv2d SSIM
  * gaussian_renderer/__init__.py:28-138            render_*
  * eval.py:122-139                                  MPJPE
The rasteriser underneath is either
  backend="ref"    : the UNMODIFIED reference CUDA kernels (oracle/_ref, GPU only), or
  backend="oracle" : the C restatement (oracle/rast_oracle.c, CPU).
torch.optim.Adam is the container's torch 2.11 (reference pinned 2.5.1; algorithm
unchanged; "parity unpinned" for Adam).  Only tests/, smoke() and
bench.py's reference / cpu_baseline legs may import this module.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import rast as crast

# --------------------------------------------------------------------------- rasteriser back-ends


class _OracleRasterize(torch.autograd.Function):
    """CPU stand-in of _RasterizeGaussians backed by the C oracle."""

    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, rs):
        feats = sh.reshape(sh.shape[0], -1) if sh.numel() else colors_precomp
        fwd = crast.forward(means3D.detach().numpy(), scales.detach().numpy(), rotations.detach().numpy(),
                            opacities.detach().numpy(), feats.detach().numpy(), rs.viewmatrix.numpy(),
                            rs.projmatrix.numpy(), rs.image_width, rs.image_height, rs.tanfovx, rs.tanfovy,
                            rs.scale_modifier, None, rs.antialiasing)
        ctx.fwd, ctx.rs = fwd, rs
        ctx.save_for_backward(means3D, scales, rotations, feats)
        radii = torch.from_numpy(fwd["radii"])
        ctx.mark_non_differentiable(radii)
        return torch.from_numpy(fwd["color"]), radii, torch.from_numpy(fwd["invdepth"])

    @staticmethod
    def backward(ctx, grad_color, _, grad_depth):
        means3D, scales, rotations, feats = ctx.saved_tensors
        rs = ctx.rs
        g = crast.backward(ctx.fwd, means3D.detach().numpy(), scales.detach().numpy(), rotations.detach().numpy(),
                           feats.detach().numpy(), rs.viewmatrix.numpy(), rs.projmatrix.numpy(), rs.image_width,
                           rs.image_height, rs.tanfovx, rs.tanfovy, grad_color.numpy(),
                           None if grad_depth is None else grad_depth.numpy(), rs.scale_modifier)
        t = torch.from_numpy
        gsh = t(g["dL_dcolors"]).reshape(feats.shape[0], 1, -1)
        return (t(g["dL_dmeans3D"]), t(g["dL_dmeans2D"]), gsh, None, t(g["dL_dopacity"]),
                t(g["dL_dscales"]), t(g["dL_drotations"]), None, None)


def _rasterizer(backend, variant, raster_settings):
    if backend == "ref":
        from . import ref_rasterizer
        return ref_rasterizer.GaussianRasterizer(raster_settings, variant)

    def call(means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None, cov3D_precomp=None):
        e = torch.Tensor([])
        return _OracleRasterize.apply(means3D, means2D, shs if shs is not None else e,
                                      colors_precomp if colors_precomp is not None else e, opacities,
                                      scales, rotations, e, raster_settings)
    return call


VARIANT_OF = {"diff-gaussian-rasterization-h36m": "h36m", "diff-gaussian-rasterization-panoptic": "panoptic",
              "diff-gaussian-rasterization-op": "op"}

# --------------------------------------------------------------------------- model (scene/gaussian_model.py)


def inverse_sigmoid(x):
    return torch.log(x / (1 - x))


def get_expon_lr_func(lr_init, lr_final, lr_delay_steps=0, lr_delay_mult=1.0, max_steps=1000000):
    def helper(step):
        if step < 0 or (lr_init == 0.0 and lr_final == 0.0):
            return 0.0
        if lr_delay_steps > 0:
            delay_rate = lr_delay_mult + (1 - lr_delay_mult) * np.sin(0.5 * np.pi * np.clip(step / lr_delay_steps, 0, 1))
        else:
            delay_rate = 1.0
        t = np.clip(step / max_steps, 0, 1)
        log_lerp = np.exp(np.log(lr_init) * (1 - t) + np.log(lr_final) * t)
        return delay_rate * log_lerp
    return helper


class RefGaussianModel:
    """create_from_pcd (149-200) + training_setup (203-236) + update_learning_rate (238-248)."""

    def __init__(self, pose_3d, cfg, spatial_lr_scale, device):
        n = cfg.n_joints
        self.spatial_lr_scale = spatial_lr_scale
        pts = torch.tensor(np.asarray(pose_3d)).float().to(device)
        one_hot = torch.zeros(n, n, device=device).scatter_(1, torch.arange(n, device=device).unsqueeze(1), 1.0)
        features = one_hot[:, :, None]
        scales = torch.ones_like(pts) * cfg.scaling
        if cfg.name in ("h36m", "panoptic", "occlusion-person") and len(cfg.modifier_joints):
            scales[list(cfg.modifier_joints), ...] *= cfg.scaling_modifier
        rots = torch.zeros((n, 4), device=device)
        rots[:, 0] = 1
        opacities = inverse_sigmoid(1.0 * torch.ones((n, 1), dtype=torch.float, device=device))
        self._xyz = torch.nn.Parameter(pts.requires_grad_(True))
        self._features_dc = torch.nn.Parameter(features.transpose(1, 2).contiguous().requires_grad_(False))
        self._features_rest = torch.nn.Parameter(features[:, :, 1:].transpose(1, 2).contiguous().requires_grad_(False))
        self._scaling = torch.nn.Parameter(scales.requires_grad_(True))
        self._rotation = torch.nn.Parameter(rots.requires_grad_(True))
        self._opacity = torch.nn.Parameter(opacities.requires_grad_(cfg.opacity_on))
        groups = [
            {'params': [self._xyz], 'lr': cfg.position_lr_init * spatial_lr_scale, "name": "xyz"},
            {'params': [self._features_dc], 'lr': cfg.feature_lr, "name": "f_dc"},
            {'params': [self._features_rest], 'lr': cfg.feature_lr / 20.0, "name": "f_rest"},
            {'params': [self._opacity], 'lr': cfg.opacity_lr, "name": "opacity"},
            {'params': [self._scaling], 'lr': cfg.scaling_lr, "name": "scaling"},
            {'params': [self._rotation], 'lr': cfg.rotation_lr, "name": "rotation"},
        ]
        self.optimizer = torch.optim.Adam(groups, lr=0.0, eps=1e-15)
        self.xyz_scheduler_args = get_expon_lr_func(lr_init=cfg.position_lr_init * spatial_lr_scale,
                                                    lr_final=cfg.position_lr_final * spatial_lr_scale,
                                                    lr_delay_mult=cfg.position_lr_delay_mult,
                                                    max_steps=cfg.position_lr_max_steps)

    get_xyz = property(lambda s: s._xyz)
    get_scaling = property(lambda s: torch.exp(s._scaling))
    get_rotation = property(lambda s: F.normalize(s._rotation))
    get_opacity = property(lambda s: torch.sigmoid(s._opacity))
    get_features = property(lambda s: s._features_dc)

    def update_learning_rate(self, iteration):
        for g in self.optimizer.param_groups:
            if g["name"] == "xyz":
                g['lr'] = self.xyz_scheduler_args(iteration)
                return g['lr']


# --------------------------------------------------------------------------- render (gaussian_renderer/__init__.py:28-138)


class TorchCamera:
    def __init__(self, cam, device):
        self.uid = cam.uid
        self.image_width, self.image_height = cam.image_width, cam.image_height
        self.FoVx, self.FoVy = cam.FoVx, cam.FoVy
        self.world_view_transform = torch.from_numpy(cam.world_view_transform).to(device)
        self.full_proj_transform = torch.from_numpy(cam.full_proj_transform).to(device)
        self.camera_center = torch.from_numpy(cam.camera_center).to(device)


def render(viewpoint_camera, pc, bg_color, backend, variant, scaling_modifier=1.0, debug=False, antialiasing=False):
    from .ref_rasterizer import GaussianRasterizationSettings
    screenspace_points = torch.zeros_like(pc.get_xyz, dtype=pc.get_xyz.dtype, requires_grad=True) + 0
    tanfovx = math.tan(viewpoint_camera.FoVx * 0.5)
    tanfovy = math.tan(viewpoint_camera.FoVy * 0.5)
    rs = GaussianRasterizationSettings(
        image_height=int(viewpoint_camera.image_height), image_width=int(viewpoint_camera.image_width),
        tanfovx=tanfovx, tanfovy=tanfovy, bg=bg_color, scale_modifier=scaling_modifier,
        viewmatrix=viewpoint_camera.world_view_transform, projmatrix=viewpoint_camera.full_proj_transform,
        sh_degree=0, campos=viewpoint_camera.camera_center, prefiltered=False, debug=debug, antialiasing=antialiasing)
    rasterizer = _rasterizer(backend, variant, rs)
    rendered_image, radii, depth_image = rasterizer(
        means3D=pc.get_xyz, means2D=screenspace_points, shs=pc.get_features, colors_precomp=None,
        opacities=pc.get_opacity, scales=pc.get_scaling, rotations=pc.get_rotation, cov3D_precomp=None)
    rendered_image = rendered_image.clamp(0, 1)
    return {"render": rendered_image, "viewspace_points": screenspace_points,
            "visibility_filter": (radii > 0).nonzero(), "radii": radii, "depth": depth_image}


# --------------------------------------------------------------------------- losses (utils/loss_utils.py)


def l1_loss(rendering, gt_heatmap, reduction='mean'):
    loss = torch.abs(rendering - gt_heatmap)
    return loss.mean() if reduction == 'mean' else (loss.sum() if reduction == 'sum' else loss)


def l2_loss_gaussian(rendering, gt_heatmap, reduction='mean'):
    mask = (gt_heatmap > 0) | (rendering > 0)
    error = (rendering - gt_heatmap) ** 2
    loss = error[mask]
    if reduction == 'mean':
        return loss.mean(), error
    elif reduction == 'sum':
        return loss.sum()
    return loss


def l1_loss_gaussian(rendering, gt_heatmap, reduction='mean'):
    mask = (gt_heatmap > 0) | (rendering > 0)
    loss = torch.abs(rendering - gt_heatmap)[mask]
    return loss.mean() if reduction == 'mean' else (loss.sum() if reduction == 'sum' else loss)


l1_loss_masked = l1_loss_gaussian  # utils/loss_utils.py:173-192 computes the same value


def l2_loss_gaussian_l1_loss_gaussian(rendering, gt_heatmap, lambda_loss=1.0, reduction='mean'):
    l2 = l2_loss_gaussian(rendering, gt_heatmap, reduction='none')
    l1 = l1_loss_gaussian(rendering, gt_heatmap, reduction='none')
    if reduction == 'mean':
        return (1.0 - lambda_loss) * l2.mean() + lambda_loss * l1.mean()
    return (1.0 - lambda_loss) * l2.sum() + lambda_loss * l1.sum()


def limb_3d_consistency_loss(xyz, limb_pairs):
    (a0, a1), (b0, b1), (c0, c1), (d0, d1) = limb_pairs
    l_arm = torch.norm(xyz[a0] - xyz[a1], dim=-1)
    r_arm = torch.norm(xyz[b0] - xyz[b1], dim=-1)
    l_leg = torch.norm(xyz[c0] - xyz[c1], dim=-1)
    r_leg = torch.norm(xyz[d0] - xyz[d1], dim=-1)
    return torch.norm(l_arm - r_arm) + torch.norm(l_leg - r_leg)


def _gaussian_window(window_size, sigma):
    g = torch.Tensor([math.exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)])
    return g / g.sum()


def ssim(img1, img2, window_size=11, size_average=True):
    """conv2d SSIM, utils/loss_utils.py:257-300 (== fused-ssim/tests/test.py:14-54)."""
    channel = img1.size(-3)
    w1 = _gaussian_window(window_size, 1.5).unsqueeze(1)
    window = w1.mm(w1.t()).float().unsqueeze(0).unsqueeze(0).expand(channel, 1, window_size, window_size).contiguous()
    window = window.to(img1.device).type_as(img1)
    p = window_size // 2
    mu1 = F.conv2d(img1, window, padding=p, groups=channel)
    mu2 = F.conv2d(img2, window, padding=p, groups=channel)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    sigma1_sq = F.conv2d(img1 * img1, window, padding=p, groups=channel) - mu1_sq
    sigma2_sq = F.conv2d(img2 * img2, window, padding=p, groups=channel) - mu2_sq
    sigma12 = F.conv2d(img1 * img2, window, padding=p, groups=channel) - mu1_mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    ssim_map = ((2 * mu1_mu2 + C1) * (2 * sigma12 + C2)) / ((mu1_sq + mu2_sq + C1) * (sigma1_sq + sigma2_sq + C2))
    return ssim_map.mean() if size_average else ssim_map


# --------------------------------------------------------------------------- GT heatmaps (utils/general_utils.py:175-304)


def generate_heatmaps_dense(gaussians, poses_2d, cams):
    """Dense restatement with scipy.ndimage.gaussian_filter standing in for cupyx's."""
    from scipy.ndimage import gaussian_filter
    dev = gaussians.get_xyz.device
    n_joints, n_views = gaussians.get_xyz.shape[0], len(cams)
    xyz = gaussians.get_xyz.detach()
    s = torch.exp(gaussians._scaling.detach())
    q = gaussians._rotation.detach()
    q = q / q.norm(dim=1, keepdim=True)
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                     2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                     2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], 1).reshape(-1, 3, 3)
    L = R @ torch.diag_embed(s)
    Vrk = L @ L.transpose(1, 2)
    view_matrix = torch.stack([c.world_view_transform.T for c in cams])
    tan_fovx = torch.stack([torch.tan(torch.tensor(c.FoVx * 0.5, device=dev)) for c in cams])
    tan_fovy = torch.stack([torch.tan(torch.tensor(c.FoVy * 0.5, device=dev)) for c in cams])
    focal_x = torch.stack([cams[i].image_width / (2.0 * tan_fovx[i]) for i in range(n_views)]).reshape(-1, 1)
    focal_y = torch.stack([cams[i].image_height / (2.0 * tan_fovy[i]) for i in range(n_views)]).reshape(-1, 1)
    hom = torch.cat([xyz, torch.ones((n_joints, 1), device=dev)], dim=1)
    t = torch.matmul(view_matrix, hom.T).transpose(1, 2)[:, :, :3].clone()
    limx, limy = (1.3 * tan_fovx).reshape(-1, 1), (1.3 * tan_fovy).reshape(-1, 1)
    txtz, tytz = t[:, :, 0] / t[:, :, 2], t[:, :, 1] / t[:, :, 2]
    t[:, :, 0] = torch.clamp(txtz, -limx, limx) * t[:, :, 2]
    t[:, :, 1] = torch.clamp(tytz, -limy, limy) * t[:, :, 2]
    J = torch.zeros((n_views, n_joints, 3, 3), device=dev)
    J[:, :, 0, 0] = focal_x / t[:, :, 2]
    J[:, :, 0, 2] = -(focal_x * t[:, :, 0]) / t[:, :, 2] ** 2
    J[:, :, 1, 1] = focal_y / t[:, :, 2]
    J[:, :, 1, 2] = -(focal_y * t[:, :, 1]) / t[:, :, 2] ** 2
    T = view_matrix[:, :3, :3].unsqueeze(1) @ J
    cov = T.permute(0, 1, 3, 2) @ Vrk.permute(0, 2, 1) @ T
    cov_x, cov_y, cov_z = cov[:, :, 0, 0] + 0.3, cov[:, :, 0, 1], cov[:, :, 1, 1] + 0.3
    det = cov_x * cov_z - cov_y * cov_y
    mid = 0.5 * (cov_x + cov_z)
    lambda1 = torch.sqrt(mid + torch.sqrt(torch.max(torch.tensor(0.1, device=dev), mid * mid - det)))
    lambda2 = torch.sqrt(mid - torch.sqrt(torch.max(torch.tensor(0.1, device=dev), mid * mid - det)))
    heatmaps = {}
    p2d = torch.as_tensor(np.asarray(poses_2d))
    for i_cam in range(n_views):
        H, W = cams[i_cam].image_height, cams[i_cam].image_width
        hm = np.zeros((n_joints, H, W), np.float32)
        xs = torch.clamp(p2d[i_cam, :, 0].long(), 0, W - 1)
        ys = torch.clamp(p2d[i_cam, :, 1].long(), 0, H - 1)
        for j in range(n_joints):
            hm[j, ys[j], xs[j]] = 255
            hm[j] = gaussian_filter(hm[j], sigma=[lambda1[i_cam, j].item(), lambda2[i_cam, j].item()])
        hm = torch.from_numpy(hm)
        cmin = hm.view(n_joints, -1).min(dim=-1)[0].unsqueeze(-1).unsqueeze(-1)
        cmax = hm.view(n_joints, -1).max(dim=-1)[0].unsqueeze(-1).unsqueeze(-1)
        heatmaps[str(i_cam)] = ((hm - cmin) / (cmax - cmin + 1e-8)).to(dev)
    return heatmaps


# --------------------------------------------------------------------------- the loop (train.py:74-233)


def optimise_frame(frame, cams, cfg, spatial_lr_scale, heatmaps_dense, backend="ref", device="cuda",
                   iterations=None, trace=None, init_override=None):
    """One frame of train.py's frame loop.  heatmaps_dense: list of [J,H,W] tensors (one per
    view).  Returns the final xyz [J,3] (float32 numpy).  ``trace``: optional list that receives
    per-optimiser-step dicts (params after the step) for trajectory comparisons."""
    variant = VARIANT_OF[cfg.rendering]
    iterations = cfg.iterations if iterations is None else iterations
    gaussians = RefGaussianModel(frame.pose_3d_init, cfg, spatial_lr_scale, device)
    if init_override is not None:       # tests only: start from a non-degenerate (anisotropic, rotated) state
        with torch.no_grad():
            gaussians._scaling.copy_(torch.as_tensor(init_override[0]).to(device))
            gaussians._rotation.copy_(torch.as_tensor(init_override[1]).to(device))
    tcams = [TorchCamera(c, device) for c in cams]
    bg = torch.tensor([0, 0, 0], dtype=torch.float32, device=device)
    n_views = len(tcams)
    accumulated_grads = torch.zeros((n_views,) + tuple(gaussians.get_xyz.shape), device=device)
    cam_idx_counter = 0
    for iteration in range(1, iterations + 1):
        gaussians.update_learning_rate(iteration)
        idx = cam_idx_counter % n_views
        cam_idx_counter += 1
        pkg = render(tcams[idx], gaussians, bg, backend, variant)
        image = pkg["render"]
        l2_loss, _ = l2_loss_gaussian(image, heatmaps_dense[idx], reduction="mean")
        loss = l2_loss + limb_3d_consistency_loss(gaussians.get_xyz, cfg.limb_pairs) * cfg.lambda_consistency
        params = [gaussians.get_xyz, gaussians._scaling, gaussians._rotation, gaussians._opacity]
        if not cfg.opacity_on:
            params = params[:3]
        grads = torch.autograd.grad(loss, params)
        accumulated_grads[idx, ...] = grads[0]
        gaussians._scaling.grad = grads[1]
        gaussians._rotation.grad = grads[2]
        if cfg.opacity_on:
            gaussians._opacity.grad = grads[3]
        if iteration % cfg.accumulation_steps == 0:
            gaussians._xyz.grad = accumulated_grads.mean(dim=0)
            with torch.no_grad():
                gaussians.optimizer.step()
                gaussians.optimizer.zero_grad(set_to_none=True)
            if trace is not None:
                trace.append(dict(iteration=iteration, xyz=gaussians._xyz.detach().cpu().numpy().copy(),
                                  scaling=gaussians._scaling.detach().cpu().numpy().copy(),
                                  rotation=gaussians._rotation.detach().cpu().numpy().copy(),
                                  loss=float(loss.item())))
    return gaussians._xyz.detach().cpu().numpy().copy()


def mpjpe(pred, gt):
    """eval.py:122-123: mean over joints of the Euclidean distance (mm)."""
    return float(np.linalg.norm(np.asarray(pred) - np.asarray(gt), axis=-1).mean())
