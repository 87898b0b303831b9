"""The fused per-frame optimiser against the restated train.py loop: on the C oracle (CPU, short runs), on the
golden final poses produced by the UNMODIFIED reference kernels (500 iterations), and through size-independent
properties at BASELINE.json's full sizes (determinism, batch invariance, frame independence).

Tolerance (north-star: final joints / MPJPE within 0.1 mm): MPJPE always within 0.1 mm; per-joint positions within
0.1 mm wherever the reference itself is reproducible to that level (occlusion-person: rotation lr 0).  For the
configs whose rotation group is driven by Adam with eps=1e-15 on noise-level gradients the reference's OWN two-run
spread (its backward uses unordered fp32 atomics) is 0.2-0.7 mm per joint (stored in the fixtures), so the
per-joint bound there is a small multiple of that measured spread."""
import numpy as np
import pytest
import torch

from oracle import pipeline as opipe
from skelsplat_b200 import configs, synthetic, heatmaps, trainer
from skelsplat_b200.cameras import cameras_extent
from tests.util import small_config, golden_path, have_golden

pytestmark = pytest.mark.gpu
DEV = "cuda"


def oracle_loop(cfg, seq, fi, iterations, trace=None):
    fr = seq.frames[fi]
    _, scal0, rot0, _ = trainer.initial_raw_state(cfg, fr.pose_3d_init[None])
    rois = heatmaps.generate_heatmap_rois(fr.pose_3d_init, fr.poses_2d, seq.cameras, scal0[0], rot0[0])
    dense = [torch.from_numpy(heatmaps.rois_to_dense(rois, v)) for v in range(cfg.nviews)]
    return opipe.optimise_frame(fr, seq.cameras, cfg, cameras_extent(seq.cameras), dense, backend="oracle", device="cpu",
                                iterations=iterations, trace=trace)


@pytest.mark.parametrize("name", ["h36m", "h36m-occ", "panoptic", "occlusion-person", "occlusion-person-8v"])
def test_short_run_matches_the_oracle_loop(name):
    """24 iterations = 6 Adam steps (covers the stale/zero gradient slots of the 8-view rig)."""
    cfg = small_config(configs.get_config(name))
    seq = synthetic.make_sequence(cfg, 2, seed=12)
    mine = trainer.optimize_sequence(seq, DEV, iterations=24)
    for fi in range(2):
        ref = oracle_loop(cfg, seq, fi, 24)
        moved = np.linalg.norm(ref - seq.frames[fi].pose_3d_init, axis=-1).max()
        assert moved > 1.0                                    # the optimiser actually moved the joints
        # 0.1 mm: at this reduced image scale a splat covers a handful of pixels, so one pixel flipping the alpha>=1/255
        # test between glibc's and CUDA's expf moves a gradient by ~1 %
        assert np.linalg.norm(mine[fi] - ref, axis=-1).max() < 0.1, name
        assert np.median(np.linalg.norm(mine[fi] - ref, axis=-1)) < 0.01, name


@pytest.mark.parametrize("name", ["h36m", "h36m-occ", "panoptic", "occlusion-person", "occlusion-person-8v"])
def test_full_run_against_reference_golden(name):
    if not have_golden(f"opt_{name}.npz"):
        pytest.skip("golden fixture missing")
    G = np.load(golden_path(f"opt_{name}.npz"))
    cfg = configs.get_config(name)
    seq = synthetic.make_sequence(cfg, int(G["n_frames"]), seed=int(G["seed"]))
    assert np.allclose(np.stack([f.pose_3d_init for f in seq.frames]), G["init_xyz"])      # same synthetic inputs
    mine = trainer.optimize_sequence(seq, DEV, iterations=int(G["iterations"]))
    ref, gt = G["ref_xyz"], G["gt_xyz"]
    assert abs(trainer.mpjpe(mine, gt) - trainer.mpjpe(ref, gt)) < 0.1                     # MPJPE within 0.1 mm
    dev = np.linalg.norm(mine - ref, axis=-1)
    rr = G["ref_xyz_reruns"]                                                               # the reference vs itself (3 runs of the first frames)
    spread = np.linalg.norm(rr - ref[:rr.shape[0], None], axis=-1)
    # Per joint.  The reference is NOT reproducible (unordered fp32 atomics in its backward, amplified by Adam with eps=1e-15):
    # on 64 frames its own run-to-run deviation has a heavy tail (h36m: median 0.02 mm, 99th percentile 1.8 mm, max 2.4 mm on
    # the 8 re-run frames; h36m-occ max 7.6 mm).  "Within 0.1 mm of the reference" can therefore only be asserted where the
    # reference agrees with itself to that level; elsewhere the fused result must be statistically indistinguishable from a
    # reference re-run: same median, same 99th percentile (within 2x), and the MPJPE -- the quantity the pipeline reports -- within 0.1 mm.
    assert np.median(dev) <= max(0.01, 2.0 * np.median(spread)), (name, np.median(dev), np.median(spread))
    assert np.percentile(dev, 99) <= max(0.1, 2.0 * np.percentile(spread, 99)), (name, np.percentile(dev, 99), np.percentile(spread, 99))
    if name == "occlusion-person":                 # rotation lr 0: the reference reproduces itself to 0.01 mm, so does the fused kernel
        assert dev.max() < 0.1, (name, dev.max())


@pytest.mark.parametrize("name", ["h36m", "panoptic"])
def test_two_adam_steps_from_a_non_degenerate_state(name):
    """Anisotropic scales and rotated quaternions make every gradient (incl. rotation, whose gradient is exactly zero in the
    symmetric initial state) well defined; after 2 Adam steps all raw parameters must match the oracle loop."""
    cfg = small_config(configs.get_config(name), factor=2)
    seq = synthetic.make_sequence(cfg, 1, seed=14)
    J = cfg.n_joints
    rng = np.random.default_rng(5)
    scal = (cfg.scaling + 0.8 + rng.uniform(-0.4, 0.4, (J, 3))).astype(np.float32)
    rot = rng.normal(size=(J, 4)).astype(np.float32)
    poses_init = np.stack([f.pose_3d_init for f in seq.frames]); poses_2d = np.stack([f.poses_2d for f in seq.frames])
    ps = trainer.pack_sequence(cfg, seq.cameras, poses_init, poses_2d, DEV)
    ps.scaling.copy_(torch.from_numpy(scal)[None]); ps.rotation.copy_(torch.from_numpy(rot)[None])
    trainer.optimize_packed(ps, iterations=8)
    fr = seq.frames[0]
    _, scal0, rot0, _ = trainer.initial_raw_state(cfg, fr.pose_3d_init[None])
    rois = heatmaps.generate_heatmap_rois(fr.pose_3d_init, fr.poses_2d, seq.cameras, scal0[0], rot0[0])
    dense = [torch.from_numpy(heatmaps.rois_to_dense(rois, v)) for v in range(cfg.nviews)]
    trace = []
    ref_xyz = opipe.optimise_frame(fr, seq.cameras, cfg, cameras_extent(seq.cameras), dense, backend="oracle", device="cpu",
                                   iterations=8, trace=trace, init_override=(scal, rot))
    lr_xyz = cfg.position_lr_init * cameras_extent(seq.cameras)
    assert np.abs(ps.xyz[0].cpu().numpy() - ref_xyz).max() < 0.02 * lr_xyz
    assert np.abs(ps.scaling[0].cpu().numpy() - trace[-1]["scaling"]).max() < 0.02 * cfg.scaling_lr
    assert np.abs(ps.rotation[0].cpu().numpy() - trace[-1]["rotation"]).max() < 0.02 * cfg.rotation_lr
    assert np.abs(trace[-1]["rotation"] - rot).max() > 0.5 * cfg.rotation_lr       # the rotation group really stepped


def test_full_size_properties():
    cfg = configs.H36M
    seq = synthetic.make_sequence(cfg, 6, seed=21)
    a = trainer.optimize_sequence(seq, DEV)
    b = trainer.optimize_sequence(seq, DEV)
    assert np.array_equal(a, b)                                                            # deterministic (no atomics)
    sub = synthetic.Sequence(cfg=cfg, cameras=seq.cameras, frames=[seq.frames[4], seq.frames[1]])
    c = trainer.optimize_sequence(sub, DEV)
    assert np.array_equal(c[0], a[4]) and np.array_equal(c[1], a[1])                       # frames are independent
    assert np.isfinite(a).all()
    gt = np.stack([f.pose_3d_gt for f in seq.frames]); init = np.stack([f.pose_3d_init for f in seq.frames])
    assert trainer.mpjpe(a, gt) < trainer.mpjpe(init, gt) + 2.0


def test_loss_decreases_and_is_reported():
    cfg = configs.H36M_OCC
    seq = synthetic.make_sequence(cfg, 4, seed=3)
    poses_init = np.stack([f.pose_3d_init for f in seq.frames]); poses_2d = np.stack([f.poses_2d for f in seq.frames])
    host = trainer.pack_host(cfg, seq.cameras, poses_init, poses_2d)
    first = trainer.optimize_packed(trainer.pack_sequence(cfg, seq.cameras, poses_init, poses_2d, DEV, host=host), iterations=4)[1].cpu().numpy()
    last = trainer.optimize_packed(trainer.pack_sequence(cfg, seq.cameras, poses_init, poses_2d, DEV, host=host))[1].cpu().numpy()
    assert (last < first).all() and (last > 0).all()


def test_capacity_overflow_is_reported():
    cfg = configs.PANOPTIC
    seq = synthetic.make_sequence(cfg, 1, seed=0)
    with pytest.raises(Exception, match="r_capacity"):
        trainer.optimize_sequence(seq, DEV, iterations=8, r_capacity=64)


def test_unsupported_loss_fails_loudly():
    from dataclasses import replace
    cfg = replace(configs.H36M, loss_function="l1")
    seq = synthetic.make_sequence(cfg, 1, seed=0)
    with pytest.raises(NotImplementedError):
        trainer.optimize_sequence(seq, DEV, iterations=4)


def test_streaming_pipeline_equals_resident_run():
    """trainer.StreamingOptimizer (pinned host batches, double-buffered copies on side streams) returns exactly what the
    resident path returns, batch after batch, including a batch shorter in ROI data than the buffers."""
    cfg = configs.H36M
    seqs = [synthetic.make_sequence(cfg, 8, seed=30 + i) for i in range(3)]
    hosts, refs = [], []
    for sq in seqs:
        pi = np.stack([f.pose_3d_init for f in sq.frames]); p2 = np.stack([f.poses_2d for f in sq.frames])
        h = trainer.pack_host(cfg, seqs[0].cameras, pi, p2)
        hosts.append({k: torch.from_numpy(v).pin_memory() for k, v in h.items()})
        ps = trainer.pack_sequence(cfg, seqs[0].cameras, pi, p2, DEV, host=h)
        refs.append(trainer.optimize_packed(ps, iterations=60)[0].cpu().numpy())
    cap = max(int(h["roi_data"].numel()) for h in hosts) + 1000
    so = trainer.StreamingOptimizer(cfg, seqs[0].cameras, 8, cap, DEV, iterations=60)
    tickets = [so.submit(h) for h in hosts[:2]]
    out0 = so.result(tickets[0])
    tickets.append(so.submit(hosts[2]))
    outs = [out0, so.result(tickets[1]), so.result(tickets[2])]
    for a, b in zip(outs, refs):
        assert np.array_equal(a, b)
    assert so.launches == 3


def test_capacity_retry_is_transparent_and_exact():
    """Frames whose (Gaussian,tile) lists outgrow the default capacity are re-run from their initial state with a doubled
    capacity; the result equals a run that had the larger capacity from the start."""
    from dataclasses import replace
    cfg = replace(configs.H36M, scaling=3.7)                       # larger splats: > 340 pairs per view > default 320
    seq = synthetic.make_sequence(cfg, 3, seed=17)
    assert trainer.default_r_capacity(cfg) == 320
    with pytest.raises(Exception, match="r_capacity"):
        trainer.optimize_sequence(seq, DEV, iterations=40, r_capacity=256)
    auto = trainer.optimize_sequence(seq, DEV, iterations=40)       # default capacity + automatic retry
    big = trainer.optimize_sequence(seq, DEV, iterations=40, r_capacity=512)
    assert np.array_equal(auto, big)


def test_streaming_retries_only_the_overflowed_frames():
    """A streamed batch in which SOME frames outgrow the capacity: the status words flag them, result() re-runs just those
    frames from their initial state (ROI buffers of the slot still intact) and returns what the resident path returns."""
    from dataclasses import replace
    cfg = replace(configs.H36M, scaling=3.7)
    big = synthetic.make_sequence(cfg, 3, seed=17)
    small = synthetic.make_sequence(configs.H36M, 3, seed=18)
    pi = np.concatenate([np.stack([f.pose_3d_init for f in s.frames]) for s in (big, small)])
    p2 = np.concatenate([np.stack([f.poses_2d for f in s.frames]) for s in (big, small)])
    # frames 3..5 keep the default splat size: overwrite their raw scales after packing
    h = trainer.pack_host(cfg, big.cameras, pi, p2)
    h2 = trainer.pack_host(configs.H36M, big.cameras, pi[3:], p2[3:])
    n_big = int(h["roi_offset"][3].min())
    h["scaling"][3:] = h2["scaling"]; h["roi_rect"][3:] = h2["roi_rect"]; h["roi_offset"][3:] = h2["roi_offset"] + n_big
    h["roi_data"] = np.concatenate([h["roi_data"][:n_big], h2["roi_data"]])
    ps = trainer.pack_sequence(cfg, big.cameras, pi, p2, DEV, host=h)
    st = trainer._launch(ps, trainer.make_opt_config(cfg, 320, 40), trainer.xyz_lr_table(cfg, ps.spatial_lr_scale, 40),
                         torch.empty(6, device=DEV)).cpu().numpy()
    assert st[:3].any() and not st[3:].any()                       # the premise: only the big-splat frames overflow
    ps = trainer.pack_sequence(cfg, big.cameras, pi, p2, DEV, host=h)
    ref = trainer.optimize_packed(ps, iterations=40)[0].cpu().numpy()
    so = trainer.StreamingOptimizer(cfg, big.cameras, 6, int(h["roi_data"].size), DEV, iterations=40)
    out = so.result(so.submit({k: torch.from_numpy(v).pin_memory() for k, v in h.items()}))
    assert np.array_equal(out, ref)
    assert so.launches == 1


def _dense_binning(cfg, ps, frame, view, r_capacity=2048):
    """Binning state of the dense-contract op (raster_dense.cu, bit-exact against the reference kernels in test_gpu_raster.py)
    for one frame/view of a PackedSequence at its CURRENT parameters."""
    from skelsplat_b200 import rasterizer as R
    J = cfg.n_joints
    dims = ps.dims.cpu().numpy(); tf = ps.tanfov.cpu().numpy()
    W, H = int(dims[view, 0]), int(dims[view, 1])
    means = ps.xyz[frame:frame + 1].contiguous()
    scales = torch.exp(ps.scaling[frame:frame + 1]).contiguous()
    rots = torch.nn.functional.normalize(ps.rotation[frame:frame + 1], dim=-1).contiguous()
    opac = torch.sigmoid(ps.opacity[frame:frame + 1]).contiguous()
    feats = torch.eye(J, device=DEV)
    _, radii, _, st = R.rasterize_batched(means, scales, rots, opac, feats, ps.viewmatrix[view].reshape(1, 4, 4), ps.projmatrix[view].reshape(1, 4, 4),
                                          W, H, float(tf[view, 0]), float(tf[view, 1]), r_capacity=r_capacity, render_invdepth=False)
    return st.parse(0), (means, scales, rots, opac, W, H, float(tf[view, 0]), float(tf[view, 1]))


@pytest.mark.parametrize("name", ["h36m", "h36m-occ", "panoptic", "occlusion-person", "occlusion-person-8v"])
def test_fused_kernel_binning_is_bit_exact(name):
    """North-star level 1 on the kernel that sets the headline: the fused optimiser's OWN tile lists (closed-form sorted
    positions, optimizer.cu phase B), dumped through ssb_optimize_frames_debug, equal the dense op's point_list / active tiles /
    tile ranges bit for bit -- at step 0 (initial state) and at a later Adam step (the state the kernel itself produced), for
    every slot of the step group incl. the 8-view rig's alternating views -- and, where oracle/_ref is present, the UNMODIFIED
    reference kernels' sorted point list and ranges."""
    from oracle import ref_rasterizer as refr
    cfg = configs.get_config(name)
    seq = synthetic.make_sequence(cfg, 3, seed=41)
    poses_init = np.stack([f.pose_3d_init for f in seq.frames]); poses_2d = np.stack([f.poses_2d for f in seq.frames])
    host = trainer.pack_host(cfg, seq.cameras, poses_init, poses_2d)
    variant = opipe.VARIANT_OF[cfg.rendering]
    for step, frame in ((0, 0), (0, 2), (7, 1)):
        ps = trainer.pack_sequence(cfg, seq.cameras, poses_init, poses_2d, DEV, host=host)
        if step:
            trainer.optimize_packed(ps, iterations=step * cfg.accumulation_steps)       # parameters at the start of Adam step `step`
        state_at_step = [t.clone() for t in (ps.xyz, ps.scaling, ps.rotation, ps.opacity)]
        ps0 = trainer.pack_sequence(cfg, seq.cameras, poses_init, poses_2d, DEV, host=host)
        slots = trainer.debug_binning(ps0, frame=frame, step=step)
        assert len(slots) == cfg.accumulation_steps
        for k, sl in enumerate(slots):
            assert sl["view"] == (step * cfg.accumulation_steps + k) % cfg.nviews and sl["status"] == 0
            for dst, src in zip((ps.xyz, ps.scaling, ps.rotation, ps.opacity), state_at_step):
                dst.copy_(src)
            dn, (means, scales, rots, opac, W, H, tfx, tfy) = _dense_binning(cfg, ps, frame, sl["view"])
            assert sl["R"] == dn["R"] and sl["n_active"] == dn["n_active"] and sl["R"] > 0
            assert np.array_equal(sl["point_list"], dn["point_list"])
            assert np.array_equal(sl["inv_pos"], dn["inv_pos"])
            assert np.array_equal(sl["tile_ids"], dn["tile_ids"])
            assert np.array_equal(sl["tile_starts"], dn["tile_ranges"][:, 0])
            ends = np.concatenate([sl["tile_starts"][1:], [sl["R"]]])
            assert np.array_equal(ends, dn["tile_ranges"][:, 1])
            if refr.available(variant):
                J = cfg.n_joints
                e = torch.Tensor([]); bg = torch.zeros(32, device=DEV)
                cam = seq.cameras[sl["view"]]
                tt = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
                Rn, _, _, geom, binning, img, _ = refr.rasterize_forward(
                    variant, bg, means[0], e, opac[0].reshape(-1, 1), scales[0], rots[0], 1.0, e, tt(cam.world_view_transform),
                    tt(cam.full_proj_transform), tfx, tfy, H, W, torch.eye(J, device=DEV).reshape(J, 1, J), 0, tt(cam.camera_center))
                rs = refr.RefState(geom, binning, img, Rn, J, W, H, variant).parse()
                assert Rn == sl["R"]
                assert np.array_equal(rs["point_list"], sl["point_list"])
                dense_ranges = np.zeros_like(rs["ranges"])
                dense_ranges[sl["tile_ids"], 0] = sl["tile_starts"]; dense_ranges[sl["tile_ids"], 1] = ends
                assert np.array_equal(rs["ranges"], dense_ranges)


def test_frames_beyond_the_fused_capacity_fall_back_to_the_dense_path():
    """A frame whose (Gaussian,tile) lists exceed the fused kernel's 1024-pair ceiling (here: very large splats) no longer costs
    the batch its results: it is optimised through the dense drop-in path (train.py's loop on the dense op) and equals a
    direct run of that path; the other frames of the batch are untouched by it."""
    from dataclasses import replace
    from skelsplat_b200.training import optimise_frame_dropin
    big_cfg = replace(configs.H36M, scaling=4.6)
    big = synthetic.make_sequence(big_cfg, 1, seed=23)
    small = synthetic.make_sequence(configs.H36M, 2, seed=24)
    pi = np.concatenate([np.stack([f.pose_3d_init for f in s.frames]) for s in (big, small)])
    p2 = np.concatenate([np.stack([f.poses_2d for f in s.frames]) for s in (big, small)])
    h = trainer.pack_host(big_cfg, big.cameras, pi, p2)
    h2 = trainer.pack_host(configs.H36M, big.cameras, pi[1:], p2[1:])
    n_big = int(h["roi_offset"][1].min())
    h["scaling"][1:] = h2["scaling"]; h["roi_rect"][1:] = h2["roi_rect"]; h["roi_offset"][1:] = h2["roi_offset"] + n_big
    h["roi_data"] = np.concatenate([h["roi_data"][:n_big], h2["roi_data"]])
    iters = 8
    ps = trainer.pack_sequence(big_cfg, big.cameras, pi, p2, DEV, host=h)
    st = trainer._launch(ps, trainer.make_opt_config(big_cfg, 1024, iters), trainer.xyz_lr_table(big_cfg, ps.spatial_lr_scale, iters),
                         torch.empty(3, device=DEV)).cpu().numpy()
    assert st[0] != 0 and not st[1:].any()                        # the premise: frame 0 outgrows even the maximum capacity
    ps = trainer.pack_sequence(big_cfg, big.cameras, pi, p2, DEV, host=h)
    out = trainer.optimize_packed(ps, iterations=iters)[0].cpu().numpy()
    assert np.isfinite(out).all()
    ps2 = trainer.pack_sequence(big_cfg, big.cameras, pi, p2, DEV, host=h)
    direct = optimise_frame_dropin(big.frames[0], big.cameras, big_cfg, heatmaps_dense=trainer.dense_roi_heatmaps(ps2, 0), device=DEV, iterations=iters)
    assert np.linalg.norm(out[0] - direct, axis=-1).max() < 1e-3
    assert np.linalg.norm(out[0] - pi[0], axis=-1).max() > 0.5    # it was optimised
    ref = trainer.optimize_sequence(synthetic.Sequence(cfg=configs.H36M, cameras=big.cameras, frames=small.frames), DEV, iterations=iters)
    assert np.array_equal(out[1:], ref)


def test_dropin_op_capacity_overflow_is_loud():
    """with debug=False the dense op used to return a plausible but wrong image when a view needed more than
    r_capacity pairs.  Now the image is NaN-filled on the device and backward raises (no per-render host sync)."""
    from skelsplat_b200 import rasterizer as R
    from skelsplat_b200.gaussian_renderer import render_functions
    from skelsplat_b200.gaussian_model import GaussianModel
    from skelsplat_b200.training import TorchCamera
    from types import SimpleNamespace
    cfg = configs.H36M
    seq = synthetic.make_sequence(cfg, 1, seed=2)
    gm = GaussianModel(1, "default", DEV)
    gm.create_from_pcd(np.asarray(seq.frames[0].pose_3d_init, np.float32), seq.cameras, cameras_extent(seq.cameras), True, 3.0, 17, 1.0, "h36m")
    pipe = SimpleNamespace(debug=False, antialiasing=False, compute_cov3D_python=False, convert_SHs_python=False)
    bg = torch.zeros(3, device=DEV)
    saved = R.DEFAULT_R_CAPACITY
    try:
        R.DEFAULT_R_CAPACITY = 64
        pkg = render_functions[cfg.rendering](TorchCamera(seq.cameras[0], DEV), gm, pipe, bg)
        assert torch.isnan(pkg["render"]).any()
        with pytest.raises(Exception, match="capacity exceeded"):
            torch.autograd.grad(pkg["render"].nansum(), [gm._xyz])
        R.DEFAULT_R_CAPACITY = saved
        pkg = render_functions[cfg.rendering](TorchCamera(seq.cameras[0], DEV), gm, pipe, bg)
        assert torch.isfinite(pkg["render"]).all()
        torch.autograd.grad(pkg["render"].sum(), [gm._xyz])
    finally:
        R.DEFAULT_R_CAPACITY = saved
