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


def adam_step_table(cfg, spatial_lr_scale, iterations=None):
    """[n_steps, 5] float32: the step-dependent scalars of torch.optim.Adam's foreach path for every optimiser step of a frame --
    -lr_xyz/bc1, -lr_scaling/bc1, -lr_rotation/bc1, -lr_opacity/bc1, sqrt(bc2), bc = 1 - beta ** step -- computed with python
    floats (fp64) exactly as torch/optim/adam.py does and rounded to fp32 where torch hands them to an fp32 tensor op.  The xyz
    learning rate is the one update_learning_rate set at the stepping iteration (train.py:134, scene/gaussian_model.py:238-248)."""
    from .trainer import xyz_lr_table
    iterations = cfg.iterations if iterations is None else iterations
    lr = xyz_lr_table(cfg, spatial_lr_scale, iterations)
    n_steps = iterations // cfg.accumulation_steps
    tab = np.zeros((max(n_steps, 1), 5), np.float32)
    beta1, beta2 = 0.9, 0.999
    for s in range(n_steps):
        step = s + 1
        bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
        it = step * cfg.accumulation_steps
        tab[s] = [(float(lr[it]) / bc1) * -1, (cfg.scaling_lr / bc1) * -1, (cfg.rotation_lr / bc1) * -1, (cfg.opacity_lr / bc1) * -1, bc2 ** 0.5]
    return tab


class GraphedFrameOptimizer:
    """train.py's per-frame loop on the DENSE drop-in surface -- the rasteriser autograd op, the fused loss kernels, torch
    autograd for the activations -- captured ONCE per camera rig in CUDA graphs and replayed for every frame.

    One graph holds a whole Adam step: the ``accumulation_steps`` iteration bodies (render one view -> l2_gaussian + limb
    consistency -> autograd.grad -> gradient bookkeeping, train.py:136-179) followed by the optimiser step as one kernel
    (``ssb_adam_frame_step``: torch.optim.Adam's arithmetic with the per-step host scalars in a device table and the step
    counter on the device, which is what makes it replayable).  A frame is then ``iterations / accumulation_steps`` graph
    launches with no other host work; parameters, Adam moments, gradient slots and the GT heatmaps live in static buffers that
    ``optimise`` re-initialises in place.  Differences from ``optimise_frame_dropin`` that do not change a result:
    ``clamp(0, 1)`` after the render is omitted (the op's output lies in [0, 0.99]; clamp and its gradient are the identity
    there) and the dense ``error`` map train.py:150 unpacks but never reads is not written.
    This is the path for what the fused kernel (trainer.optimize_sequence) does not cover -- other dense losses, frames beyond
    its 1024-pair capacity -- and the honest upper bound of "swap the packages, keep the dense loop"."""

    def __init__(self, cfg, cams, device="cuda", iterations=None):
        import math
        from . import lib as _L
        from .gaussian_model import GaussianModel as GM
        self.cfg, self.cams, self.device = cfg, cams, device
        self.iterations = cfg.iterations if iterations is None else iterations
        self.n_steps = self.iterations // cfg.accumulation_steps
        J, V = cfg.n_joints, len(cams)
        self.extent = cameras_extent(cams)
        self.model = GM(1, "default", device)
        self.model.create_from_pcd(np.zeros((J, 3), np.float32), cams, self.extent, cfg.opacity_on, cfg.scaling, J, cfg.scaling_modifier, cfg.name)
        self.init_scaling = self.model._scaling.detach().clone(); self.init_rotation = self.model._rotation.detach().clone()
        self.init_opacity = self.model._opacity.detach().clone()
        self.tcams = [TorchCamera(c, device) for c in cams]
        self.heatmaps = [torch.zeros((J, c.image_height, c.image_width), dtype=torch.float32, device=device) for c in cams]
        self._rects = None
        self.accumulated_grads = torch.zeros((V, J, 3), dtype=torch.float32, device=device)
        self.static_g = [torch.zeros_like(p) for p in (self.model._scaling, self.model._rotation, self.model._opacity)]
        self.exp_avg = torch.zeros(11 * J, dtype=torch.float32, device=device); self.exp_avg_sq = torch.zeros_like(self.exp_avg)
        self.step_counter = torch.zeros(1, dtype=torch.int32, device=device)
        self.table = torch.from_numpy(adam_step_table(cfg, self.extent, self.iterations)).to(device)
        self.pipe = SimpleNamespace(debug=False, antialiasing=cfg.antialiasing, compute_cov3D_python=False, convert_SHs_python=False)
        self.bg = torch.tensor([0, 0, 0], dtype=torch.float32, device=device)
        self.data_root = "data/" + cfg.name
        if cfg.loss_function not in losses:
            raise NotImplementedError(f"loss_function={cfg.loss_function!r} is not implemented on the dense surface")
        self._L = _L
        acc = cfg.accumulation_steps
        self.n_patterns = (acc * V // math.gcd(acc, V)) // acc          # step groups with distinct view sets (1 for V | acc; 2 for the 8-view rig)
        self.graphs = None

    # ---- one iteration body (train.py:136-179) on static buffers
    def _body(self, idx):
        from . import loss_utils as LU
        import importlib
        import math
        m, cam, cfg = self.model, self.tcams[idx], self.cfg
        rast = importlib.import_module(cfg.rendering.replace("-", "_"))          # the variant package (NUM_CHANNELS) the config's render_* binds
        settings = rast.GaussianRasterizationSettings(
            image_height=int(cam.image_height), image_width=int(cam.image_width), tanfovx=math.tan(cam.FoVx * 0.5), tanfovy=math.tan(cam.FoVy * 0.5),
            bg=self.bg, scale_modifier=1.0, viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform, sh_degree=m.active_sh_degree,
            campos=cam.camera_center, prefiltered=False, debug=False, antialiasing=cfg.antialiasing)
        screenspace = torch.zeros_like(m.get_xyz, requires_grad=True) + 0
        image, _, _ = rast.GaussianRasterizer(raster_settings=settings)(
            means3D=m.get_xyz, means2D=screenspace, shs=m.get_features, colors_precomp=None, opacities=m.get_opacity, scales=m.get_scaling,
            rotations=m.get_rotation, cov3D_precomp=None)
        crit = losses[cfg.loss_function]
        out = crit(image, self.heatmaps[idx], None, cfg.lambda_loss_function, reduction="mean", want_error=False) \
            if crit is LU.l2_loss_gaussian else crit(image, self.heatmaps[idx], None, cfg.lambda_loss_function, reduction="mean")
        l2 = out[0] if isinstance(out, tuple) else out
        loss = l2 + consistency_losses[cfg.consistency_loss](m.get_xyz, self.data_root, reduction="mean") * cfg.lambda_consistency
        grads = torch.autograd.grad(loss, [m.get_xyz, m._scaling, m._rotation, m._opacity])
        self.accumulated_grads[idx].copy_(grads[0])
        for dst, g in zip(self.static_g, grads[1:]):
            dst.copy_(g)

    def _adam(self):
        import ctypes as C
        L, m, J = self._L.lib(), self.model, self.cfg.n_joints
        p = self._L.ptr
        self._L.check(L.ssb_adam_frame_step(C.c_int(J), C.c_int(len(self.cams)), p(m._xyz), p(m._scaling), p(m._rotation), p(m._opacity),
                                            p(self.accumulated_grads), p(self.static_g[0]), p(self.static_g[1]), p(self.static_g[2]), p(self.exp_avg),
                                            p(self.exp_avg_sq), p(self.table), C.c_int(self.n_steps), p(self.step_counter),
                                            C.c_float(float(np.float32(1 - 0.9))), C.c_float(0.999), C.c_float(float(np.float32(1 - 0.999))),
                                            C.c_float(1e-15), self._L.current_stream()), "ssb_adam_frame_step")

    def _step_group(self, pattern):
        acc, V = self.cfg.accumulation_steps, len(self.cams)
        for k in range(acc):
            self._body((pattern * acc + k) % V)
        self._adam()

    def capture(self):
        """Warm up on a side stream, then capture one graph per distinct step group (shared memory pool: they never overlap)."""
        dev = self.device
        self._reset(np.zeros((self.cfg.n_joints, 3), np.float32) + np.array([0.0, 0.0, 1000.0], np.float32))
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for pat in range(self.n_patterns):
                self._step_group(pat)
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graphs, pool = [], None
        for pat in range(self.n_patterns):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pool):
                self._step_group(pat)
            pool = g.pool()
            self.graphs.append(g)

    def _reset(self, pose_init):
        m = self.model
        with torch.no_grad():
            m._xyz.copy_(torch.as_tensor(np.asarray(pose_init, np.float32)).to(self.device))
            m._scaling.copy_(self.init_scaling); m._rotation.copy_(self.init_rotation); m._opacity.copy_(self.init_opacity)
        for t in (self.accumulated_grads, self.exp_avg, self.exp_avg_sq, self.step_counter, *self.static_g):
            t.zero_()

    def load_heatmaps(self, dense=None, rois=None):
        """GT heatmaps into the static buffers: ``dense`` = list of [J,H,W] tensors, or ``rois`` = heatmaps.HeatmapROIs (host) /
        (rect [V,J,4] numpy, offset [V,J] numpy, data device tensor) -- factored patches, scattered on the GPU."""
        if dense is not None:
            for dst, src in zip(self.heatmaps, dense):
                dst.copy_(torch.as_tensor(src).to(self.device))
            self._rects = None
            return
        rect, offset, data = (rois.rect, rois.offset, torch.from_numpy(rois.data).to(self.device)) if hasattr(rois, "rect") else rois
        if self._rects is None:
            for hm in self.heatmaps:
                hm.zero_()
        else:
            for v, hm in enumerate(self.heatmaps):                 # clear only the previous frame's windows
                for j, (x0, y0, w, h) in enumerate(self._rects[v]):
                    hm[j, y0:y0 + h, x0:x0 + w] = 0
        for v, hm in enumerate(self.heatmaps):
            for j in range(self.cfg.n_joints):
                x0, y0, w, h = (int(a) for a in rect[v, j]); o = int(offset[v, j])
                hm[j, y0:y0 + h, x0:x0 + w] = data[o:o + h, None] * data[None, o + h:o + h + w]
        self._rects = [[tuple(int(a) for a in rect[v, j]) for j in range(self.cfg.n_joints)] for v in range(len(self.cams))]

    def optimise(self, pose_init, dense=None, rois=None):
        """One frame: final xyz [J,3] (float32 numpy).  Give the GT heatmaps as ``dense`` or ``rois`` (see load_heatmaps)."""
        if self.graphs is None:
            self.capture()
        self.load_heatmaps(dense, rois)
        self._reset(pose_init)
        for s in range(self.n_steps):
            self.graphs[s % self.n_patterns].replay()
        out = self.model._xyz.detach().cpu().numpy().copy()
        if not np.isfinite(out).all():
            raise self._L.SkelSplatLibraryError("non-finite poses from the graphed dense loop (rasteriser capacity exceeded? see rasterizer.DEFAULT_R_CAPACITY)")
        return out
