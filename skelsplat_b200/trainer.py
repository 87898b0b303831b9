"""Batched, fused per-frame optimisation: the B200-native form of train.py's loop.

``optimize_sequence`` is what replaces ``training()`` (train.py:56-244) for a whole
sequence: every frame's 500-iteration optimisation runs inside ONE persistent CUDA
kernel (csrc/optimizer.cu) -- no per-iteration launches, host syncs, dense images or
per-frame disk I/O.  Semantics kept from the reference: one view per
iteration round-robin, Adam step every ``accumulation_steps`` iterations on the mean of
the per-view xyz gradient slots (stale/zero slots included), scaling/rotation/opacity
gradients from the group's last view, xyz learning rate taken at the stepping iteration,
``l2_gaussian`` + 1e-5 * limb consistency.
"""
import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import lib as _L
from .cameras import cameras_extent
from .configs import SceneConfig
from .heatmaps import generate_heatmap_rois


def expon_lr(step, lr_init, lr_final, lr_delay_steps=0, lr_delay_mult=1.0, max_steps=1000000):
    """get_expon_lr_func(...)(step), utils/general_utils.py:38-71, in fp64 like the reference."""
    if step < 0 or (lr_init == 0.0 and lr_final == 0.0):
        return 0.0
    if lr_delay_steps > 0:
        delay_rate = lr_delay_mult + (1 - lr_delay_mult) * np.sin(0.5 * np.pi * np.clip(step / lr_delay_steps, 0, 1))
    else:
        delay_rate = 1.0
    t = np.clip(step / max_steps, 0, 1)
    return float(delay_rate * np.exp(np.log(lr_init) * (1 - t) + np.log(lr_final) * t))


def xyz_lr_table(cfg: SceneConfig, spatial_lr_scale: float, iterations: Optional[int] = None):
    """lr of the xyz group at iteration i (update_learning_rate, scene/gaussian_model.py:238-248)."""
    n = cfg.iterations if iterations is None else iterations
    return np.array([expon_lr(i, cfg.position_lr_init * spatial_lr_scale, cfg.position_lr_final * spatial_lr_scale,
                              lr_delay_mult=cfg.position_lr_delay_mult, max_steps=cfg.position_lr_max_steps)
                     for i in range(n + 1)], np.float64)


def initial_raw_state(cfg: SceneConfig, poses_init: np.ndarray):
    """create_from_pcd (scene/gaussian_model.py:149-200) for F frames: raw (pre-activation) parameters."""
    F, J = poses_init.shape[0], cfg.n_joints
    xyz = poses_init.astype(np.float32)
    scaling = np.full((F, J, 3), cfg.scaling, np.float32)
    if cfg.scaling > 0.0 and len(cfg.modifier_joints):
        scaling[:, list(cfg.modifier_joints), :] *= np.float32(cfg.scaling_modifier)
    rotation = np.zeros((F, J, 4), np.float32)
    rotation[..., 0] = 1
    opacity = np.full((F, J), np.inf, np.float32)       # inverse_sigmoid(1.0)
    return xyz, scaling, rotation, opacity


@dataclass
class PackedSequence:
    """Device-resident inputs of the fused optimiser for F frames of one camera rig."""
    cfg: SceneConfig
    n_frames: int
    xyz: torch.Tensor            # [F,J,3]
    scaling: torch.Tensor        # [F,J,3] raw (log) scale
    rotation: torch.Tensor       # [F,J,4] raw quaternion
    opacity: torch.Tensor        # [F,J]   logit
    viewmatrix: torch.Tensor     # [V,16]
    projmatrix: torch.Tensor     # [V,16]
    dims: torch.Tensor           # [V,2] int32
    tanfov: torch.Tensor         # [V,2]
    roi_rect: torch.Tensor       # [F,V,J,4] int32
    roi_offset: torch.Tensor     # [F,V,J] int64
    roi_data: torch.Tensor       # [total] float32
    spatial_lr_scale: float
    Wmax: int
    Hmax: int


def pack_host(cfg: SceneConfig, cams, poses_init, poses_2d):
    """Host-side (numpy) packing: initial state + GT heatmap ROIs for F frames.  Returns a dict of
    numpy arrays (pinned-memory friendly) -- per-frame setup, timed separately from the hot loop."""
    F = poses_init.shape[0]
    xyz, scaling, rotation, opacity = initial_raw_state(cfg, poses_init)
    V, J = len(cams), cfg.n_joints
    rects = np.zeros((F, V, J, 4), np.int32)
    offs = np.zeros((F, V, J), np.int64)
    chunks, total = [], 0
    for f in range(F):
        rois = generate_heatmap_rois(poses_init[f], poses_2d[f], cams, scaling[f], rotation[f])
        rects[f] = rois.rect
        offs[f] = rois.offset + total
        chunks.append(rois.data)
        total += rois.data.size
    data = np.concatenate(chunks) if chunks else np.zeros(0, np.float32)
    return dict(xyz=xyz, scaling=scaling, rotation=rotation, opacity=opacity, roi_rect=rects, roi_offset=offs, roi_data=data)


def camera_tensors(cams, device):
    vm = torch.from_numpy(np.stack([c.world_view_transform.reshape(16) for c in cams])).to(device)
    pm = torch.from_numpy(np.stack([c.full_proj_transform.reshape(16) for c in cams])).to(device)
    dims = torch.tensor([[c.image_width, c.image_height] for c in cams], dtype=torch.int32, device=device)
    tanfov = torch.tensor([[c.tanfovx, c.tanfovy] for c in cams], dtype=torch.float32, device=device)
    return vm, pm, dims, tanfov


def pack_sequence(cfg: SceneConfig, cams, poses_init, poses_2d, device="cuda", host=None) -> PackedSequence:
    """Upload a host-prepared batch (pack_host: numpy heatmap ROIs, the readable specification of the setup and the route for
    callers that already hold heatmaps on the host).  The production route prepares everything on the GPU instead:
    setup_gpu.pack_sequence_gpu / StreamingOptimizer.submit_detections.  Either way the optimisation itself has no CPU path."""
    host = pack_host(cfg, cams, poses_init, poses_2d) if host is None else host
    vm, pm, dims, tanfov = camera_tensors(cams, device)
    t = lambda a: torch.from_numpy(a).to(device, non_blocking=True)
    return PackedSequence(cfg=cfg, n_frames=poses_init.shape[0], xyz=t(host["xyz"]), scaling=t(host["scaling"]),
                          rotation=t(host["rotation"]), opacity=t(host["opacity"]), viewmatrix=vm, projmatrix=pm,
                          dims=dims, tanfov=tanfov, roi_rect=t(host["roi_rect"]), roi_offset=t(host["roi_offset"]),
                          roi_data=t(host["roi_data"]), spatial_lr_scale=cameras_extent(cams),
                          Wmax=max(c.image_width for c in cams), Hmax=max(c.image_height for c in cams))


def make_opt_config(cfg: SceneConfig, r_capacity=256, iterations=None):
    if cfg.loss_function != "l2_gaussian":
        # train.py:150 unpacks `l2_loss, error = opt_criterion(...)`; l2_gaussian is the only entry of the reference's
        # loss table that returns that tuple (utils/loss_utils.py:86-100), i.e. the only one its training loop can run.
        # The other losses stay available on the dense surface (skelsplat_b200.loss_utils / ssb_loss_forward).
        raise NotImplementedError(f"fused optimiser implements loss_function='l2_gaussian' (got {cfg.loss_function!r}); "
                                  "use skelsplat_b200.training.optimise_frame_dropin for the dense losses")
    oc = _L.OptConfig()
    oc.J, oc.V = cfg.n_joints, cfg.nviews
    oc.iterations = cfg.iterations if iterations is None else iterations
    oc.accumulation_steps = cfg.accumulation_steps
    oc.lambda_consistency = cfg.lambda_consistency if cfg.consistency_loss != "none" else 0.0
    flat = [i for pair in cfg.limb_pairs for i in pair]
    for i in range(8):
        oc.limb_pairs[i] = flat[i]
    oc.lr_scaling, oc.lr_rotation, oc.lr_opacity = cfg.scaling_lr, cfg.rotation_lr, cfg.opacity_lr
    oc.beta1, oc.beta2, oc.eps = 0.9, 0.999, 1e-15      # torch.optim.Adam(l, lr=0.0, eps=1e-15), gaussian_model.py:217-218
    oc.r_capacity = r_capacity
    oc.antialiasing = int(cfg.antialiasing)
    oc.max_unrolled_list = int(os.environ.get("SKELSPLAT_B200_UNROLL", tuned_max_unrolled_list(cfg)))
    oc.resident_record_slots = int(os.environ.get("SKELSPLAT_B200_RECORD_SLOTS", 0))      # 0 = auto (developer knob; results are bit-identical)
    return oc


def tuned_max_unrolled_list(cfg: SceneConfig):
    """Measured on B200 (scripts/gpu_tune_opt.py): the unrolled path for tile lists of 5 Gaussians is +6 % on H36M and +12 % on
    Panoptic shapes, -7 % on the 8-view Occlusion-Person shape (many short-lived 5-lists of small splats)."""
    return 4 if cfg.name.startswith("occlusion-person") else 5


def default_r_capacity(cfg: SceneConfig):
    """(Gaussian,tile) pairs per view held in shared memory.  ~10/joint at H36M/OP scale, 25-60/joint at Panoptic scale.
    Small capacities raise occupancy; frames that outgrow the capacity are detected on the device and re-run (below)."""
    if cfg.name.startswith("panoptic"):
        return 896        # 0 of 4 096 synthetic frames outgrow 768; 896 leaves margin and, vs 1 024, more L1 for the heatmap profiles (+2 %)
    if cfg.name.startswith("occlusion-person"):
        return 512
    return 320        # H36M: 256 is outgrown by ~1 % of the synthetic frames during optimisation; 320 costs 0.5 % throughput


MAX_R_CAPACITY = 1024


def _launch(ps: PackedSequence, oc, lr, final_loss):
    L = _L.lib()
    cfg = ps.cfg
    lr_c = (C.c_double * len(lr))(*lr.tolist())
    cams = _L.Cameras(cfg.nviews, _L.ptr(ps.viewmatrix), _L.ptr(ps.projmatrix), _L.ptr(ps.dims), _L.ptr(ps.tanfov),
                      ps.Wmax, ps.Hmax, 0.0, 0.0, int(cfg.antialiasing))
    F = ps.xyz.shape[0]
    ws = torch.zeros(max(int(L.ssb_optimize_workspace_bytes(C.byref(oc), C.c_int(F))), 4) // 4, dtype=torch.int32, device=ps.xyz.device)
    rc = L.ssb_optimize_frames(C.byref(oc), C.c_int(F), C.byref(cams), lr_c, _L.ptr(ps.xyz), _L.ptr(ps.scaling),
                               _L.ptr(ps.rotation), _L.ptr(ps.opacity), _L.ptr(ps.roi_rect), _L.ptr(ps.roi_offset),
                               _L.ptr(ps.roi_data), _L.ptr(final_loss), _L.ptr(ws), _L.current_stream())
    _L.check(rc, "ssb_optimize_frames")
    return ws[:F]


def debug_binning(ps: PackedSequence, frame=0, step=0, r_capacity=None, iterations=None):
    """Debug accessor (ssb_optimize_frames_debug): run the fused optimiser on a COPY of ``ps`` for ``iterations`` (default:
    just far enough to reach Adam step ``step``) and return the kernel's own binning state of ``frame`` at that step, one dict
    per slot (= iteration of the step group): view, R, point_list [R] (sorted Gaussian ids), inv_pos [R], tile_ids [n_active]
    (row-major tile index y * grid_x + x), tile_starts [n_active].  This is what the parity tests compare bit-for-bit with
    the dense op's / the reference's binningState (rasterizer_impl.cu:70-138, 303-320)."""
    L = _L.lib()
    cfg = ps.cfg
    rcap = default_r_capacity(cfg) if r_capacity is None else r_capacity
    acc = cfg.accumulation_steps
    iters = (step + 1) * acc if iterations is None else iterations
    oc = make_opt_config(cfg, rcap, iters)
    lr = xyz_lr_table(cfg, ps.spatial_lr_scale, oc.iterations)
    lr_c = (C.c_double * len(lr))(*lr.tolist())
    cams = _L.Cameras(cfg.nviews, _L.ptr(ps.viewmatrix), _L.ptr(ps.projmatrix), _L.ptr(ps.dims), _L.ptr(ps.tanfov),
                      ps.Wmax, ps.Hmax, 0.0, 0.0, int(cfg.antialiasing))
    F = ps.xyz.shape[0]
    st = [t.clone() for t in (ps.xyz, ps.scaling, ps.rotation, ps.opacity)]
    ws = torch.zeros(max(int(L.ssb_optimize_workspace_bytes(C.byref(oc), C.c_int(F))), 4) // 4, dtype=torch.int32, device=ps.xyz.device)
    dbg = torch.full((acc, 4 + 4 * rcap), -1, dtype=torch.int32, device=ps.xyz.device)
    rc = L.ssb_optimize_frames_debug(C.byref(oc), C.c_int(F), C.byref(cams), lr_c, _L.ptr(st[0]), _L.ptr(st[1]), _L.ptr(st[2]),
                                     _L.ptr(st[3]), _L.ptr(ps.roi_rect), _L.ptr(ps.roi_offset), _L.ptr(ps.roi_data), None, _L.ptr(ws),
                                     C.c_int(frame), C.c_int(step), _L.ptr(dbg), _L.current_stream())
    _L.check(rc, "ssb_optimize_frames_debug")
    d = dbg.cpu().numpy()
    dims = ps.dims.cpu().numpy()
    out = []
    for k in range(acc):
        R, nact, view, status = (int(x) for x in d[k, :4])
        body = d[k, 4:].reshape(4, rcap)
        gx = (int(dims[view, 0]) + 15) // 16
        tile = body[2, :nact]
        out.append(dict(view=view, R=R, n_active=nact, status=status, point_list=body[0, :R].astype(np.uint32),
                        inv_pos=body[1, :R].astype(np.uint32), tile_ids=((tile >> 8) * gx + (tile & 255)).astype(np.uint32),
                        tile_starts=body[3, :nact].astype(np.uint32)))
    return out


def optimize_packed(ps: PackedSequence, iterations=None, r_capacity=None, final_loss=None, check=True):
    """Run the fused optimiser in place on a PackedSequence.  Returns (xyz [F,J,3], final_loss [F]).

    check=True (default) reads the per-frame status words back (one small D2H, a host sync) and transparently
    re-runs, from their initial state and with a doubled capacity, the frames whose (Gaussian,tile) lists outgrew
    ``r_capacity``; check=False skips the read-back (benchmark inner loops that verify separately)."""
    cfg = ps.cfg
    if cfg.loss_function != "l2_gaussian":
        raise NotImplementedError("the fused optimiser implements the loss every shipped config uses (l2_gaussian); "
                                  "other losses run through the drop-in per-iteration path")
    rcap = default_r_capacity(cfg) if r_capacity is None else r_capacity
    oc = make_opt_config(cfg, rcap, iterations)
    lr = xyz_lr_table(cfg, ps.spatial_lr_scale, oc.iterations)
    F = ps.n_frames
    if final_loss is None:
        final_loss = torch.empty(F, dtype=torch.float32, device=ps.xyz.device)
    init = tuple(t.clone() for t in (ps.xyz, ps.scaling, ps.rotation, ps.opacity)) if check else None
    status = _launch(ps, oc, lr, final_loss)
    if check:
        bad = torch.nonzero(status != 0).flatten()
        if bad.numel():
            if r_capacity is not None:
                raise _L.SkelSplatLibraryError(f"{bad.numel()} frame(s) exceeded r_capacity={rcap} (Gaussian,tile) pairs per view")
            retry_overflowed(ps, bad, tuple(t[bad] for t in init), rcap, iterations, final_loss)
    return ps.xyz, final_loss


def dense_roi_heatmaps(ps: PackedSequence, frame):
    """The frame's GT heatmaps as the dense [J,H,W] tensors of the reference contract, scattered from the ROI patches on the GPU."""
    cfg = ps.cfg
    dims = ps.dims.cpu().numpy(); rect = ps.roi_rect[frame].cpu().numpy(); off = ps.roi_offset[frame].cpu().numpy()
    out = []
    for v in range(cfg.nviews):
        hm = torch.zeros((cfg.n_joints, int(dims[v, 1]), int(dims[v, 0])), dtype=torch.float32, device=ps.xyz.device)
        for j in range(cfg.n_joints):
            x0, y0, w, h = (int(a) for a in rect[v, j])
            o = int(off[v, j])      # factored patch: col[h] | row[w]; the heatmap value is their fp32 product (heatmaps.HeatmapROIs)
            hm[j, y0:y0 + h, x0:x0 + w] = ps.roi_data[o:o + h, None] * ps.roi_data[None, o + h:o + h + w]
        out.append(hm)
    return out


def dense_fallback(ps: PackedSequence, frames, init, iterations=None, final_loss=None):
    """Frames whose (Gaussian,tile) lists outgrow the fused kernel's shared-memory ceiling (MAX_R_CAPACITY pairs per view: a
    close-up, or splats grown very large) are optimised through the dense drop-in surface instead -- train.py's own loop on
    the dense rasteriser op (capacity 16 384 pairs), the fused loss kernels and torch Adam: the same algorithm, ~10^4 x slower
    per frame, so one such frame does not cost the batch its results.  ``init`` = (xyz, scaling, rotation, opacity) of those
    frames.  Returns the indices (into ``frames``) that could not be optimised even so (their poses are set to NaN)."""
    from types import SimpleNamespace
    from . import rasterizer as _R
    from .training import optimise_frame_dropin
    cfg = ps.cfg
    cams = [SimpleNamespace(uid=v) for v in range(cfg.nviews)]
    vm, pm = ps.viewmatrix.cpu().numpy(), ps.projmatrix.cpu().numpy()
    dims, tf = ps.dims.cpu().numpy(), ps.tanfov.cpu().numpy()
    import math
    for v, c in enumerate(cams):
        c.image_width, c.image_height = int(dims[v, 0]), int(dims[v, 1])
        c.FoVx, c.FoVy = 2.0 * math.atan(float(tf[v, 0])), 2.0 * math.atan(float(tf[v, 1]))
        c.world_view_transform, c.full_proj_transform = vm[v].reshape(4, 4), pm[v].reshape(4, 4)
        c.camera_center = np.linalg.inv(vm[v].reshape(4, 4).T)[:3, 3].astype(np.float32)
    failed = []
    saved = _R.DEFAULT_R_CAPACITY
    _R.DEFAULT_R_CAPACITY = _R.MAX_R_CAPACITY
    try:
        for n, f in enumerate(frames.tolist()):
            fr = SimpleNamespace(pose_3d_init=init[0][n].cpu().numpy(), poses_2d=None)
            try:
                st = optimise_frame_dropin(fr, cams, cfg, heatmaps_dense=dense_roi_heatmaps(ps, f), device=ps.xyz.device, iterations=iterations,
                                           init_state=(init[1][n], init[2][n]), return_state=True, spatial_lr_scale=ps.spatial_lr_scale)
                if not torch.isfinite(st[0]).all():
                    raise _L.SkelSplatLibraryError("non-finite result (dense op capacity exceeded)")
                ps.xyz[f] = st[0]; ps.scaling[f] = st[1]; ps.rotation[f] = st[2]; ps.opacity[f] = st[3].reshape(-1)
            except _L.SkelSplatLibraryError:
                ps.xyz[f] = float("nan")
                failed.append(n)
            if final_loss is not None:
                final_loss[f] = float("nan")          # the dense loop does not report it
    finally:
        _R.DEFAULT_R_CAPACITY = saved
    return failed


def retry_overflowed(ps: PackedSequence, bad, init_bad, rcap, iterations=None, final_loss=None):
    """Re-run ONLY the frames ``bad`` (indices into ps) from their initial state ``init_bad`` = (xyz, scaling, rotation, opacity)
    with the capacity doubled until they fit; results are written into ps in place.  Exact: a frame's result does not depend
    on the capacity it ran with, nor on the other frames of the launch.  Frames that outgrow even MAX_R_CAPACITY go through
    ``dense_fallback``; only if that fails too is an error raised -- after every other frame's result is in place, naming the
    failed frame indices (``err.failed_frames``)."""
    cfg = ps.cfg
    lr = xyz_lr_table(cfg, ps.spatial_lr_scale, cfg.iterations if iterations is None else iterations)
    cur = tuple(t.contiguous() for t in init_bad)
    while bad.numel():
        if rcap >= MAX_R_CAPACITY:
            failed = dense_fallback(ps, bad, cur, iterations, final_loss)
            if failed:
                err = _L.SkelSplatLibraryError(f"frame(s) {[int(bad[i]) for i in failed]} exceeded r_capacity={rcap} (Gaussian,tile) pairs per view "
                                               "in the fused optimiser and the dense fallback's capacity too; every other frame's result is valid")
                err.failed_frames = [int(bad[i]) for i in failed]
                raise err
            return
        rcap = min(2 * rcap, MAX_R_CAPACITY)
        oc = make_opt_config(cfg, rcap, iterations)
        sub = PackedSequence(cfg=cfg, n_frames=int(bad.numel()), xyz=cur[0].clone(), scaling=cur[1].clone(), rotation=cur[2].clone(),
                             opacity=cur[3].clone(), viewmatrix=ps.viewmatrix, projmatrix=ps.projmatrix, dims=ps.dims, tanfov=ps.tanfov,
                             roi_rect=ps.roi_rect[bad].contiguous(), roi_offset=ps.roi_offset[bad].contiguous(), roi_data=ps.roi_data,
                             spatial_lr_scale=ps.spatial_lr_scale, Wmax=ps.Wmax, Hmax=ps.Hmax)
        sub_loss = torch.empty(sub.n_frames, dtype=torch.float32, device=ps.xyz.device)
        sub_status = _launch(sub, oc, lr, sub_loss)
        good = sub_status == 0
        idx = bad[good]
        ps.xyz[idx] = sub.xyz[good]; ps.scaling[idx] = sub.scaling[good]; ps.rotation[idx] = sub.rotation[good]
        ps.opacity[idx] = sub.opacity[good]
        if final_loss is not None:
            final_loss[idx] = sub_loss[good]
        bad = bad[~good]
        cur = tuple(t[~good].contiguous() for t in cur)


def optimize_sequence(seq, device="cuda", iterations=None, r_capacity=None):
    """Whole synthetic Sequence -> final poses [F,J,3] (numpy, float32)."""
    poses_init = np.stack([f.pose_3d_init for f in seq.frames])
    poses_2d = np.stack([f.poses_2d for f in seq.frames])
    ps = pack_sequence(seq.cfg, seq.cameras, poses_init, poses_2d, device)
    xyz, _ = optimize_packed(ps, iterations, r_capacity)
    return xyz.cpu().numpy()


def mpjpe(pred, gt):
    """eval.py:122-123: mean over joints (and frames) of ||pred - gt||_2, millimetres."""
    return float(np.linalg.norm(np.asarray(pred, np.float64) - np.asarray(gt, np.float64), axis=-1).mean())


class StreamingOptimizer:
    """End-to-end pipeline for sequences that live in HOST memory: batch i+1's host->device copy (initial state + GT ROIs,
    from pinned buffers) overlaps batch i's optimisation on a second CUDA stream; final poses and the per-frame status words
    come back device->host per batch.  Double-buffered: at most two batches are resident.

        so = StreamingOptimizer(cfg, cams, frames_per_batch, roi_floats_per_batch)
        t0 = so.submit(host_batch0); t1 = so.submit(host_batch1)      # dicts of pinned tensors as trainer.pack_host() lays them out
        xyz0 = so.result(t0)                                           # blocks on batch 0 only
    """

    def __init__(self, cfg: SceneConfig, cams, frames_per_batch, roi_floats, device="cuda", iterations=None, r_capacity=None):
        self.cfg, self.cams, self.F, self.device = cfg, cams, frames_per_batch, device
        self.iterations, self.r_capacity = iterations, r_capacity
        J, V = cfg.n_joints, cfg.nviews
        vm, pm, dims, tanfov = camera_tensors(cams, device)
        ext = cameras_extent(cams)
        Wmax, Hmax = max(c.image_width for c in cams), max(c.image_height for c in cams)
        mk = lambda shape, dt: torch.empty(shape, dtype=dt, device=device)
        self.slots = []
        for _ in range(2):
            ps = PackedSequence(cfg=cfg, n_frames=frames_per_batch, xyz=mk((frames_per_batch, J, 3), torch.float32),
                                scaling=mk((frames_per_batch, J, 3), torch.float32), rotation=mk((frames_per_batch, J, 4), torch.float32),
                                opacity=mk((frames_per_batch, J), torch.float32), viewmatrix=vm, projmatrix=pm, dims=dims, tanfov=tanfov,
                                roi_rect=mk((frames_per_batch, V, J, 4), torch.int32), roi_offset=mk((frames_per_batch, V, J), torch.int64),
                                roi_data=mk((roi_floats,), torch.float32), spatial_lr_scale=ext, Wmax=Wmax, Hmax=Hmax)
            self.slots.append(dict(ps=ps, out=torch.empty((frames_per_batch, J, 3), dtype=torch.float32).pin_memory(),
                                   status=torch.empty((frames_per_batch,), dtype=torch.int32).pin_memory(),
                                   loss=torch.empty(frames_per_batch, dtype=torch.float32, device=device),
                                   copied=torch.cuda.Event(), done=torch.cuda.Event(), host=None, busy=False))
        self.copy_stream = torch.cuda.Stream(device=device)
        self.compute_stream = torch.cuda.Stream(device=device)
        self.n_submitted = 0
        self.launches = 0
        self._det = None              # buffers of the detections-in path, created on first use

    # ---- detections in, poses out: initial guess + GT heatmap ROIs are produced on the GPU (setup_gpu), so a step moves
    # ---- F*V*J*2 detection floats (+ optional initial poses) over PCIe instead of the ROI patches themselves
    def _det_buffers(self):
        if self._det is None:
            F, J, V, dev = self.F, self.cfg.n_joints, self.cfg.nviews, self.device
            _, scal, rot, opa = initial_raw_state(self.cfg, np.zeros((1, J, 3), np.float32))
            tmpl = tuple(torch.from_numpy(np.ascontiguousarray(np.broadcast_to(a, (F,) + a.shape[1:]))).to(dev) for a in (scal, rot, opa))
            per = []
            for _ in range(2):
                per.append(dict(p2d=torch.empty((F, V, J, 2), dtype=torch.float32, device=dev),
                                sigma=torch.empty((F, V, J, 2), dtype=torch.float32, device=dev),
                                center=torch.empty((F, V, J, 2), dtype=torch.int32, device=dev),
                                size=torch.empty((F, V, J), dtype=torch.int64, device=dev),
                                setup_status=torch.zeros(1, dtype=torch.int32, device=dev),
                                setup_status_host=torch.zeros(1, dtype=torch.int32).pin_memory()))
            self._det = dict(tmpl=tmpl, per=per, P=torch.as_tensor(np.asarray([c.P3x4() for c in self.cams], np.float64)).to(dev))
        return self._det

    def submit_detections(self, host):
        """host: dict of PINNED tensors  poses_2d [F,V,J,2] fp32  and optionally  xyz [F,J,3] fp32 (initial guess; without it
        the DLT triangulation of the detections is used, triangulation.py:122-150).  Returns a ticket for result()."""
        from . import setup_gpu
        i = self.n_submitted
        sl = self.slots[i % 2]
        if sl["busy"]:
            sl["done"].synchronize()
        det = self._det_buffers()
        d, ps = det["per"][i % 2], sl["ps"]
        if host["poses_2d"].shape[0] != self.F:
            raise ValueError("batch does not fit the streaming buffers")
        with torch.cuda.stream(self.copy_stream):
            d["p2d"].copy_(host["poses_2d"], non_blocking=True)
            if host.get("xyz") is not None:
                ps.xyz.copy_(host["xyz"], non_blocking=True)
            sl["copied"].record(self.copy_stream)
        with torch.cuda.stream(self.compute_stream):
            self.compute_stream.wait_event(sl["copied"])
            if host.get("xyz") is None:
                ps.xyz.copy_(setup_gpu.triangulate_dlt(det["P"], d["p2d"], self.device))
            for dst, src in zip((ps.scaling, ps.rotation, ps.opacity), det["tmpl"]):
                dst.copy_(src)
            d["setup_status"].zero_()
            setup_gpu.generate_heatmap_rois_into(self.cfg, ps.viewmatrix, ps.projmatrix, ps.dims, ps.tanfov, ps.Wmax, ps.Hmax,
                                                 ps.xyz, ps.scaling, ps.rotation, d["p2d"], ps.roi_rect, d["sigma"], d["center"],
                                                 d["size"], ps.roi_offset, ps.roi_data, d["setup_status"])
            rcap = default_r_capacity(self.cfg) if self.r_capacity is None else self.r_capacity
            oc = make_opt_config(self.cfg, rcap, self.iterations)
            lr = xyz_lr_table(self.cfg, ps.spatial_lr_scale, oc.iterations)
            status = _launch(ps, oc, lr, sl["loss"])
            self.launches += 1
            sl["out"].copy_(ps.xyz, non_blocking=True)
            sl["status"].copy_(status, non_blocking=True)
            d["setup_status_host"].copy_(d["setup_status"], non_blocking=True)
            sl["done"].record(self.compute_stream)
        sl["host"], sl["busy"], sl["det"] = host, True, d
        self.n_submitted += 1
        return i

    def submit(self, host):
        i = self.n_submitted
        sl = self.slots[i % 2]
        if sl["busy"]:
            sl["done"].synchronize()            # the buffer pair of batch i-2 must have been consumed by its kernel and copies
        ps = sl["ps"]
        n_roi = host["roi_data"].numel()
        if n_roi > ps.roi_data.numel() or host["xyz"].shape[0] != self.F:
            raise ValueError("batch does not fit the streaming buffers")
        with torch.cuda.stream(self.copy_stream):
            for k, dst in (("xyz", ps.xyz), ("scaling", ps.scaling), ("rotation", ps.rotation), ("opacity", ps.opacity),
                           ("roi_rect", ps.roi_rect), ("roi_offset", ps.roi_offset)):
                dst.copy_(host[k], non_blocking=True)
            ps.roi_data[:n_roi].copy_(host["roi_data"], non_blocking=True)
            sl["copied"].record(self.copy_stream)
        with torch.cuda.stream(self.compute_stream):
            self.compute_stream.wait_event(sl["copied"])
            rcap = default_r_capacity(self.cfg) if self.r_capacity is None else self.r_capacity
            oc = make_opt_config(self.cfg, rcap, self.iterations)
            lr = xyz_lr_table(self.cfg, ps.spatial_lr_scale, oc.iterations)
            status = _launch(ps, oc, lr, sl["loss"])
            self.launches += 1
            sl["out"].copy_(ps.xyz, non_blocking=True)
            sl["status"].copy_(status, non_blocking=True)
            sl["done"].record(self.compute_stream)
        sl["host"], sl["busy"], sl["det"] = host, True, None
        self.n_submitted += 1
        return i

    def result(self, ticket):
        sl = self.slots[ticket % 2]
        sl["done"].synchronize()
        det, host, ps = sl.get("det"), sl["host"], sl["ps"]
        if det is not None and int(det["setup_status_host"][0]) != 0:
            # rare: the ROI patches outgrew the streaming buffer -> exact synchronous re-run of the batch through the
            # exactly-sized, retrying path
            from . import setup_gpu
            with torch.cuda.stream(self.compute_stream):
                full = setup_gpu.pack_sequence_gpu(self.cfg, self.cams, host["poses_2d"], host.get("xyz"), self.device)
                optimize_packed(full, self.iterations, None)
                sl["out"].copy_(full.xyz)
            self.compute_stream.synchronize()
        elif int(sl["status"].max()) != 0:
            # rare: some frames outgrew r_capacity -> exact re-run of THOSE frames from their initial state (the slot's ROI
            # buffers are still intact: the slot is not reused before this call returns)
            if self.r_capacity is not None:
                raise _L.SkelSplatLibraryError(f"frame(s) exceeded r_capacity={self.r_capacity} (Gaussian,tile) pairs per view")
            with torch.cuda.stream(self.compute_stream):
                bad = torch.nonzero(sl["status"].to(self.device) != 0).flatten()
                if det is None:
                    init = tuple(host[k].to(self.device, non_blocking=True)[bad] for k in ("xyz", "scaling", "rotation", "opacity"))
                else:
                    if host.get("xyz") is not None:
                        xyz0 = host["xyz"].to(self.device, non_blocking=True)[bad]
                    else:
                        from . import setup_gpu
                        xyz0 = setup_gpu.triangulate_dlt(self._det["P"], det["p2d"][bad], self.device).to(torch.float32)
                    init = (xyz0,) + tuple(t[bad] for t in self._det["tmpl"])
                retry_overflowed(ps, bad, init, default_r_capacity(self.cfg), self.iterations)
                sl["out"].copy_(ps.xyz)
            self.compute_stream.synchronize()
        return sl["out"].numpy().copy()

    def synchronize(self):
        self.copy_stream.synchronize()
        self.compute_stream.synchronize()
