"""GaussianModel: parameter and optimiser handling of the skeletal Gaussians (host side, torch).

Mirror of the hot-path part of the reference's scene/gaussian_model.py -- same attribute
names (``_xyz, _features_dc, _features_rest, _scaling, _rotation, _opacity``), getters
(102-143), ``create_from_pcd`` (149-200), ``training_setup`` (203-236),
``update_learning_rate`` (238-248), ``capture``/``restore`` (68-100) -- so that
``gaussian_renderer.render_*`` and a train.py-style loop work unchanged on top of it.
Densify/prune/PLY-load (341-498) are never reached by SkelSplat and are out of scope.
"""
import numpy as np
import torch
from torch import nn

from .trainer import expon_lr


def inverse_sigmoid(x):
    return torch.log(x / (1 - x))            # utils/general_utils.py:27-28


def get_expon_lr_func(lr_init, lr_final, lr_delay_steps=0, lr_delay_mult=1.0, max_steps=1000000):
    def helper(step):
        return expon_lr(step, lr_init, lr_final, lr_delay_steps, lr_delay_mult, max_steps)
    return helper


def build_rotation(r):
    """utils/general_utils.py:87-108."""
    norm = torch.sqrt(r[:, 0] * r[:, 0] + r[:, 1] * r[:, 1] + r[:, 2] * r[:, 2] + r[:, 3] * r[:, 3])
    q = r / norm[:, None]
    R = torch.zeros((q.size(0), 3, 3), device=r.device)
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - r * z); R[:, 0, 2] = 2 * (x * z + r * y)
    R[:, 1, 0] = 2 * (x * y + r * z); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - r * x)
    R[:, 2, 0] = 2 * (x * z - r * y); R[:, 2, 1] = 2 * (y * z + r * x); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def build_scaling_rotation(s, r):
    """utils/general_utils.py:110-119: L = R @ diag(s)."""
    L = torch.zeros((s.shape[0], 3, 3), dtype=torch.float, device=s.device)
    R = build_rotation(r)
    L[:, 0, 0] = s[:, 0]; L[:, 1, 1] = s[:, 1]; L[:, 2, 2] = s[:, 2]
    return R @ L


def strip_symmetric(sym):
    out = torch.zeros((sym.shape[0], 6), dtype=torch.float, device=sym.device)
    out[:, 0] = sym[:, 0, 0]; out[:, 1] = sym[:, 0, 1]; out[:, 2] = sym[:, 0, 2]
    out[:, 3] = sym[:, 1, 1]; out[:, 4] = sym[:, 1, 2]; out[:, 5] = sym[:, 2, 2]
    return out


_MODIFIER_JOINTS = {"h36m": [3, 6, 12, 13, 15, 16], "panoptic": [8, 14, 4, 5, 10, 11], "occlusion-person": [3, 6, 10, 11, 13, 14]}


class GaussianModel:
    def __init__(self, sh_degree, optimizer_type="default", device="cuda"):
        self.active_sh_degree = 0
        self.optimizer_type = optimizer_type
        self.max_sh_degree = sh_degree
        self.device = device
        e = torch.empty(0)
        self._xyz = self._features_dc = self._features_rest = self._scaling = self._rotation = self._opacity = e
        self.max_radii2D = self.xyz_gradient_accum = self.denom = e
        self.optimizer = None
        self.percent_dense = 0
        self.spatial_lr_scale = 0
        self.scaling_activation = torch.exp
        self.scaling_inverse_activation = torch.log
        self.opacity_activation = torch.sigmoid
        self.inverse_opacity_activation = inverse_sigmoid
        self.rotation_activation = torch.nn.functional.normalize

    @staticmethod
    def covariance_activation(scaling, scaling_modifier, rotation):
        L = build_scaling_rotation(scaling_modifier * scaling, rotation)
        return strip_symmetric(L @ L.transpose(1, 2))

    def capture(self):
        return (self.active_sh_degree, self._xyz, self._features_dc, self._features_rest, self._scaling, self._rotation,
                self._opacity, self.max_radii2D, self.xyz_gradient_accum, self.denom, self.optimizer.state_dict(),
                self.spatial_lr_scale)

    def restore(self, model_args, training_args):
        (self.active_sh_degree, self._xyz, self._features_dc, self._features_rest, self._scaling, self._rotation,
         self._opacity, self.max_radii2D, xyz_gradient_accum, denom, opt_dict, self.spatial_lr_scale) = model_args
        self.training_setup(training_args)
        self.xyz_gradient_accum = xyz_gradient_accum
        self.denom = denom
        self.optimizer.load_state_dict(opt_dict)

    get_scaling = property(lambda self: self.scaling_activation(self._scaling))
    get_rotation = property(lambda self: self.rotation_activation(self._rotation))
    get_xyz = property(lambda self: self._xyz)
    get_features = property(lambda self: self._features_dc)            # scene/gaussian_model.py:114-118: dc only
    get_features_dc = property(lambda self: self._features_dc)
    get_features_rest = property(lambda self: self._features_rest)
    get_opacity = property(lambda self: self.opacity_activation(self._opacity))

    def get_covariance(self, scaling_modifier=1):
        return self.covariance_activation(self.get_scaling, scaling_modifier, self._rotation)

    def oneupSHdegree(self):
        if self.active_sh_degree < self.max_sh_degree:
            self.active_sh_degree += 1

    def create_from_pcd(self, pcd, cam_infos, spatial_lr_scale, opacity_on, scaling, n_joints, scaling_modifier=1.0, scene_type="h36m"):
        """``pcd`` is anything with a ``.points`` [J,3] array (or the array itself): the initial pose."""
        dev = self.device
        self.spatial_lr_scale = spatial_lr_scale
        points = np.asarray(getattr(pcd, "points", pcd))
        fused_point_cloud = torch.tensor(points).float().to(dev)
        joint_indices = torch.arange(n_joints).unsqueeze(1).to(dev)
        one_hot = torch.zeros(n_joints, n_joints, device=dev).scatter_(1, joint_indices, 1.0)
        features = one_hot[:, :, None]
        scales = torch.from_numpy(points).float().to(dev)
        if scaling > 0.0:
            scales = torch.ones_like(scales) * scaling
            if scene_type in _MODIFIER_JOINTS:                  # exact match: "h36m-occ" gets no modifier
                scales[_MODIFIER_JOINTS[scene_type], ...] *= scaling_modifier
        rots = torch.zeros((fused_point_cloud.shape[0], 4), device=dev)
        rots[:, 0] = 1
        opacities = self.inverse_opacity_activation(1.0 * torch.ones((fused_point_cloud.shape[0], 1), dtype=torch.float, device=dev))
        self._xyz = nn.Parameter(fused_point_cloud.requires_grad_(True))
        self._features_dc = nn.Parameter(features.transpose(1, 2).contiguous().requires_grad_(False))
        self._features_rest = nn.Parameter(features[:, :, 1:].transpose(1, 2).contiguous().requires_grad_(False))
        self._scaling = nn.Parameter(scales.requires_grad_(True))
        self._rotation = nn.Parameter(rots.requires_grad_(True))
        self._opacity = nn.Parameter(opacities.requires_grad_(bool(opacity_on)))
        self.max_radii2D = torch.zeros((self.get_xyz.shape[0]), device=dev)

    def training_setup(self, training_args):
        self.percent_dense = getattr(training_args, "percent_dense", 0.01)
        self.xyz_gradient_accum = torch.zeros((self.get_xyz.shape[0], 1), device=self.device)
        self.denom = torch.zeros((self.get_xyz.shape[0], 1), device=self.device)
        groups = [
            {'params': [self._xyz], 'lr': training_args.position_lr_init * self.spatial_lr_scale, "name": "xyz"},
            {'params': [self._features_dc], 'lr': training_args.feature_lr, "name": "f_dc"},
            {'params': [self._features_rest], 'lr': training_args.feature_lr / 20.0, "name": "f_rest"},
            {'params': [self._opacity], 'lr': training_args.opacity_lr, "name": "opacity"},
            {'params': [self._scaling], 'lr': training_args.scaling_lr, "name": "scaling"},
            {'params': [self._rotation], 'lr': training_args.rotation_lr, "name": "rotation"},
        ]
        # "sparse_adam" needs a rasteriser the reference does not ship either: it falls back to Adam too (222-226)
        self.optimizer = torch.optim.Adam(groups, lr=0.0, eps=1e-15)
        self.xyz_scheduler_args = get_expon_lr_func(lr_init=training_args.position_lr_init * self.spatial_lr_scale,
                                                    lr_final=training_args.position_lr_final * self.spatial_lr_scale,
                                                    lr_delay_mult=training_args.position_lr_delay_mult,
                                                    max_steps=training_args.position_lr_max_steps)

    def update_learning_rate(self, iteration):
        for param_group in self.optimizer.param_groups:
            if param_group["name"] == "xyz":
                lr = self.xyz_scheduler_args(iteration)
                param_group['lr'] = lr
                return lr
