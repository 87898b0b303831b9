"""Host-side logic (numpy / torch-CPU): cameras, DLT, GT heatmap ROIs, schedules, configs, sharding."""
import math

import numpy as np
import pytest
import torch

from skelsplat_b200 import cameras, configs, heatmaps, synthetic, trainer, triangulation
from tests.util import small_config


def test_projection_conventions():
    """full_proj maps a world point to NDC such that ndc2Pix gives fx*x/z + cx - 0.5 and
    the homogeneous w equals view-space z; matrices are stored transposed (scene/cameras.py:94-99)."""
    seq = synthetic.make_sequence(configs.H36M, 1, seed=0)
    X = seq.frames[0].pose_3d_gt
    for cam in seq.cameras:
        hom = np.concatenate([X, np.ones((X.shape[0], 1))], 1)
        clip = hom @ cam.full_proj_transform.astype(np.float64)           # row-vector convention of the stored matrices
        view = hom @ cam.world_view_transform.astype(np.float64)
        assert np.allclose(clip[:, 3], view[:, 2], rtol=1e-6)
        ndc = clip[:, :2] / clip[:, 3:4]
        pix = ((ndc + 1.0) * np.array([cam.image_width, cam.image_height]) - 1.0) * 0.5
        uv = synthetic.project(cam, X)
        assert np.allclose(pix, uv - 0.5, atol=2e-2)
        assert np.allclose(view[:, :3], X @ cam.R_w2c.T + cam.t, atol=1e-2)
        assert math.isclose(cam.tanfovx, cam.image_width / (2 * cam.K[0, 0]), rel_tol=1e-12)
        c = -cam.R_w2c.T @ cam.t
        assert np.allclose(cam.camera_center, c, atol=1e-2)
    ext = cameras.cameras_extent(seq.cameras)
    centers = np.stack([-c.R_w2c.T @ c.t for c in seq.cameras])
    assert math.isclose(ext, 1.1 * np.linalg.norm(centers - centers.mean(0), axis=1).max(), rel_tol=1e-5)


def test_dlt_matches_reference_algorithm_and_recovers_noiseless_points():
    cfg = configs.PANOPTIC
    seq = synthetic.make_sequence(cfg, 3, seed=4)
    P_list = [c.P3x4() for c in seq.cameras]
    for fr in seq.frames:
        det = np.stack([synthetic.project(c, fr.pose_3d_gt) for c in seq.cameras])
        X = triangulation.triangulate_poses(P_list, det)
        assert np.abs(X - fr.pose_3d_gt).max() < 1e-5
        # per-joint loop exactly as triangulation.py:122-150
        for j in range(cfg.n_joints):
            A = []
            for P, x in zip(P_list, fr.poses_2d[:, j]):
                A.append(x[0] * P[2, :] - P[0, :]); A.append(x[1] * P[2, :] - P[1, :])
            _, _, Vt = np.linalg.svd(np.array(A))
            Xj = Vt[-1] / Vt[-1][3]
            assert np.allclose(Xj[:3], fr.pose_3d_init[j], atol=1e-6)


@pytest.mark.parametrize("name", ["h36m", "occlusion-person"])
def test_heatmap_rois_equal_the_dense_reference_procedure(name):
    """ROI patches == delta(255) -> scipy gaussian_filter([s1,s2], truncate 4, reflect) -> per-channel min-max,
    the restatement of utils/general_utils.py:175-304 (oracle/pipeline.generate_heatmaps_dense)."""
    from oracle import pipeline as opipe
    cfg = small_config(configs.get_config(name), factor=4)
    seq = synthetic.make_sequence(cfg, 1, seed=2)
    fr = seq.frames[0]
    fr.poses_2d[0, 0] = [1.2, 3.9]                       # near a corner: exercises 'reflect' and the clamp
    fr.poses_2d[1, 1] = [-5.0, 1e4]                      # outside the image: clamped to the border
    g = opipe.RefGaussianModel(fr.pose_3d_init, cfg, 1.0, "cpu")
    tcams = [opipe.TorchCamera(c, "cpu") for c in seq.cameras]
    dense_ref = opipe.generate_heatmaps_dense(g, fr.poses_2d, tcams)
    _, scal, rot, _ = trainer.initial_raw_state(cfg, fr.pose_3d_init[None])
    rois = heatmaps.generate_heatmap_rois(fr.pose_3d_init, fr.poses_2d, seq.cameras, scal[0], rot[0])
    for v in range(cfg.nviews):
        mine = heatmaps.rois_to_dense(rois, v)
        ref = dense_ref[str(v)].numpy()
        assert mine.shape == ref.shape
        assert np.array_equal(mine > 0, ref > 0)
        assert np.abs(mine - ref).max() < 2e-6
        assert mine.reshape(cfg.n_joints, -1).max(1).min() > 0.999


def test_heatmap_sigma_uses_the_python_side_covariance_not_ewa():
    """for the isotropic initial Gaussian the GT sigma has no perspective terms."""
    cfg = configs.H36M
    seq = synthetic.make_sequence(cfg, 1, seed=0)
    fr = seq.frames[0]
    _, scal, rot, _ = trainer.initial_raw_state(cfg, fr.pose_3d_init[None])
    s1, s2 = heatmaps.heatmap_sigmas(fr.pose_3d_init, seq.cameras, scal[0], rot[0])
    cam = seq.cameras[0]
    z = (fr.pose_3d_init @ cam.R_w2c.T + cam.t)[:, 2]
    sx = math.exp(3.0) * cam.K[0, 0] / z
    lam = sx * sx + 0.3
    # lambda = mid +- sqrt(max(0.1, ~0)) for an (almost) isotropic 2D covariance
    assert np.allclose(s1[0] ** 2, lam + math.sqrt(0.1), rtol=5e-3)
    assert np.allclose(s2[0] ** 2, lam - math.sqrt(0.1), rtol=5e-3)


def test_lr_schedule_and_initial_state():
    cfg = configs.H36M
    tab = trainer.xyz_lr_table(cfg, 5000.0)
    assert len(tab) == 501
    a, b = cfg.position_lr_init * 5000.0, cfg.position_lr_final * 5000.0
    for it in (1, 4, 250, 500):
        t = it / 4000
        assert math.isclose(tab[it], math.exp(math.log(a) * (1 - t) + math.log(b) * t), rel_tol=1e-12)
    from oracle import pipeline as opipe
    f = opipe.get_expon_lr_func(a, b, lr_delay_mult=0.0, max_steps=4000)
    assert all(tab[i] == f(i) for i in range(1, 501))
    poses = np.zeros((2, 15, 3))
    xyz, scal, rot, opa = trainer.initial_raw_state(configs.OCCLUSION_PERSON, poses)
    assert np.allclose(scal[0, [3, 6, 10, 11, 13, 14]], 3.75) and np.allclose(scal[0, [0, 1, 2]], 3.0)
    xyz, scal, rot, opa = trainer.initial_raw_state(configs.H36M_OCC, np.zeros((1, 17, 3)))
    assert np.allclose(scal, 3.0)                         # "h36m-occ" matches no branch: the 1.25 modifier is a no-op
    assert np.all(rot[..., 0] == 1) and np.all(rot[..., 1:] == 0) and np.all(np.isinf(opa))


def test_configs_match_the_shipped_yaml_values():
    h, p, o = configs.H36M, configs.PANOPTIC, configs.OCCLUSION_PERSON
    assert (h.n_joints, p.n_joints, o.n_joints) == (17, 19, 15)
    assert h.position_lr_init == 0.0005 and p.position_lr_init == 0.005 and o.position_lr_init == 0.005
    assert h.rotation_lr == 0.001 and o.rotation_lr == 0.0 and p.opacity_lr == 0.005 and h.opacity_lr == 0.0
    assert h.iterations == 500 and h.accumulation_steps == 4 and h.lambda_consistency == 1e-5
    assert h.image_sizes[0] == (1002, 1000) and p.image_sizes[0] == (1920, 1080) and o.image_sizes[0] == (1280, 720)
    assert configs.get_config("occlusion-person-8v").nviews == 8


def test_shard_bounds_partition_the_frames():
    from skelsplat_b200.distributed import shard_bounds
    for n in (0, 1, 7, 64, 100001):
        for w in (1, 2, 4, 8):
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [e - s for s, e in b]
            assert max(sizes) - min(sizes) <= 1


def test_synthetic_sequences_are_seeded_and_shaped():
    for name in ("h36m", "h36m-occ", "panoptic", "occlusion-person", "occlusion-person-8v"):
        cfg = configs.get_config(name)
        a = synthetic.make_sequence(cfg, 2, seed=9); b = synthetic.make_sequence(cfg, 2, seed=9)
        assert len(a.cameras) == cfg.nviews
        assert a.frames[0].poses_2d.shape == (cfg.nviews, cfg.n_joints, 2)
        assert np.array_equal(a.frames[1].pose_3d_init, b.frames[1].pose_3d_init)
        err = np.linalg.norm(a.frames[0].pose_3d_init - a.frames[0].pose_3d_gt, axis=1).mean()
        assert err < (200 if cfg.occluded else 60)
        (l0, l1), (r0, r1) = cfg.limb_pairs[0], cfg.limb_pairs[1]
        gt = a.frames[0].pose_3d_gt
        assert math.isclose(np.linalg.norm(gt[l0] - gt[l1]), np.linalg.norm(gt[r0] - gt[r1]), rel_tol=1e-9)


def test_fused_path_rejects_losses_the_reference_loop_cannot_run():
    """train.py:150 unpacks a (loss, error) tuple, which only l2_gaussian returns: the fused optimiser implements that loss
    and refuses the others loudly instead of silently optimising a different objective."""
    import dataclasses
    bad = dataclasses.replace(configs.H36M, loss_function="l1_gaussian")
    with pytest.raises(NotImplementedError):
        trainer.make_opt_config(bad)
    assert trainer.make_opt_config(configs.H36M).J == 17


@pytest.mark.parametrize("name,seed", [("h36m", 0), ("panoptic", 1), ("occlusion-person", 2), ("h36m", 3)])
def test_closed_form_sorted_positions_equal_the_reference_sort(name, seed):
    """The fused optimiser bins without sorting: the closed-form position of every (Gaussian, tile) pair must reproduce the
    order of the reference's stable (tile | depth) radix sort -- checked here against the C oracle's sort on random,
    overlapping, partly culled Gaussians (equal depths included)."""
    from oracle import rast
    from tests import mirror_binning as binning
    from tests.util import raster_case
    cfg = small_config(configs.get_config(name), 2)
    case = raster_case(cfg, seed=seed, n_views=1, big=(seed != 3))
    means = case["means3D"].copy()
    means[2] = means[0]                                         # identical centres: equal depth bits, order decided by id
    means[3, 2] -= 5000.0                                       # behind the camera: culled, touches no tile
    W, H = int(case["dims"][0, 0]), int(case["dims"][0, 1])
    fw = rast.forward(means, case["scales"], case["rotations"], case["opacities"], case["features"], case["viewmatrix"][0],
                      case["projmatrix"][0], W, H, float(case["tanfov"][0, 0]), float(case["tanfov"][0, 1]))
    rects = fw["rects"].astype(np.int64).copy()                 # (x0, y0, x1, y1) per Gaussian
    rects[fw["radii"] <= 0] = 0
    gs, xy, pos = binning.sorted_positions(rects, fw["depths"].view(np.uint32))
    R = fw["R"]
    assert len(pos) == R and sorted(pos.tolist()) == list(range(R))       # a permutation of 0..R-1
    gx = (W + 15) // 16
    assert np.array_equal(gs[np.argsort(pos)], fw["point_list"].astype(np.int64))
    tiles_sorted = (fw["keys_sorted"] >> np.uint64(32)).astype(np.int64)
    assert np.array_equal((xy[:, 1] * gx + xy[:, 0])[np.argsort(pos)], tiles_sorted)
    assert (fw["radii"] <= 0).any() and R > 40


def test_row_band_contains_every_contributing_pixel():
    """Exactness of the row culling: brute force over pixels with the reference's fp32 power expression (forward.cu:352-364)
    never finds a contributing pixel (power >= -5.55, the cut used with opacity <= 1) outside the band, for isotropic, elongated,
    rotated, tiny and huge splats."""
    from tests import mirror_binning as binning
    rng = np.random.default_rng(0)
    f = np.float32
    checked = 0
    for trial in range(300):
        s1, s2 = np.exp(rng.uniform(-0.5, 4.0, 2))                       # std devs 0.6 .. 55 px
        th = rng.uniform(0, np.pi)
        Rm = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
        cov = Rm @ np.diag([s1 * s1, s2 * s2]) @ Rm.T + 0.3 * np.eye(2)
        inv = np.linalg.inv(cov)
        conx, cony, conz = f(inv[0, 0]), f(inv[0, 1]), f(inv[1, 1])
        px, py = f(rng.uniform(100, 400)), f(rng.uniform(100, 400))
        rlo, rhi = binning.row_band(py, conx, cony, conz)
        ys, xs = np.mgrid[0:512, 0:512]
        dx = (px - xs.astype(f)).astype(f); dy = (py - ys.astype(f)).astype(f)
        t = (dy * (dy * conz).astype(f)).astype(f)
        t = (dx * (dx * conx).astype(f) + t).astype(f)
        power = (t * f(-0.5) - (dy * (dx * cony).astype(f)).astype(f)).astype(f)
        rows = np.nonzero(((power <= 0) & (power >= f(-5.55))).any(axis=1))[0]
        if rows.size:
            assert rlo <= rows.min() and rows.max() <= rhi, (trial, rlo, rhi, rows.min(), rows.max())
            checked += 1
            if 8 < rows.min() and rows.max() < 500:                       # and the band is tight: a few rows of slack at most
                assert rows.min() - rlo <= 3 + 0.011 * (rows.max() - rows.min()) and rhi - rows.max() <= 3 + 0.011 * (rows.max() - rows.min())
    assert checked > 250


def test_fused_one_hot_backward_equals_the_reference_backward():
    """The algebra of the fused kernel (one scalar recurrence per pixel, raw moment sums, mask size from per-pair counts,
    constants applied once) against the C oracle of the reference's forward + per-channel backward fed with the l2_gaussian
    gradient: same image, loss, N and per-Gaussian gradients."""
    from oracle import rast
    from tests import mirror_fused_math as fused_math
    from tests.util import raster_case, relerr
    cfg = small_config(configs.H36M, 8)
    case = raster_case(cfg, seed=5, n_views=1, big=True)
    J = case["means3D"].shape[0]
    feats = np.eye(J, dtype=np.float32)
    W, H = int(case["dims"][0, 0]), int(case["dims"][0, 1])
    args = (case["viewmatrix"][0], case["projmatrix"][0], W, H, float(case["tanfov"][0, 0]), float(case["tanfov"][0, 1]))
    fw = rast.forward(case["means3D"], case["scales"], case["rotations"], case["opacities"], feats, *args)
    rng = np.random.default_rng(1)
    ys, xs = np.mgrid[0:H, 0:W]
    gt = np.zeros((J, H, W), np.float32)
    for j in range(J):                                          # GT blobs near (not at) the splats, exactly zero elsewhere
        cx, cy = fw["means2D"][j] + rng.normal(0, 6, 2)
        blob = np.exp(-((xs - cx) ** 2 + (ys - cy) ** 2) / (2 * 7.0 ** 2))
        gt[j] = np.where(blob > 0.05, blob, 0).astype(np.float32)
    out = fused_math.view_iteration(fw["means2D"], fw["conic_opacity"], fw["ranges"], fw["point_list"], gt, W, H)
    assert relerr(out["render"], fw["color"]) < 1e-6
    render = fw["color"]
    mask = (gt > 0) | (render > 0)
    N = int(mask.sum())
    assert out["N"] == N and 0 < (render > 0).sum() < render.size
    loss = float((((render - gt).astype(np.float64)) ** 2)[mask].sum() / N)
    assert abs(out["loss"] - loss) < 1e-6 * loss
    dL = np.where(mask, 2.0 * (render - gt) / N, 0).astype(np.float32)
    bw = rast.backward(fw, case["means3D"], case["scales"], case["rotations"], feats, *args, dL)
    assert relerr(out["dL_dmean2D"], bw["dL_dmeans2D"][:, :2]) < 1e-4
    ref_conic = np.stack([bw["dL_dconic"][:, 0, 0], bw["dL_dconic"][:, 0, 1], bw["dL_dconic"][:, 1, 1]], 1)
    assert relerr(out["dL_dconic"], ref_conic) < 1e-4
    assert relerr(out["dL_dopacity"], bw["dL_dopacity"][:, 0]) < 1e-4
    assert np.abs(bw["dL_dmeans2D"]).max() > 0


def test_adam_mirror_matches_torch_adam():
    """The in-kernel Adam (host fp64 step sizes rounded to fp32, eps = 1e-15 added AFTER the bias-corrected sqrt) against
    torch.optim.Adam on the CPU, including the noise-level gradients that eps = 1e-15 turns into full-size steps."""
    from tests import mirror_fused_math as fused_math
    rng = np.random.default_rng(0)
    n = 64
    p0 = rng.normal(0, 100, n).astype(np.float32)
    grads = [(rng.normal(0, 1, n) * 10.0 ** rng.uniform(-12, 1, n)).astype(np.float32) for _ in range(6)]
    lrs = [2.5 * 0.97 ** s for s in range(6)]                   # the xyz group's lr changes every step (update_learning_rate)
    tp = torch.nn.Parameter(torch.from_numpy(p0.copy()))
    opt = torch.optim.Adam([{"params": [tp], "lr": 0.0}], lr=0.0, eps=1e-15)
    p, m, v = p0.copy(), np.zeros(n, np.float32), np.zeros(n, np.float32)
    for s, (g, lr) in enumerate(zip(grads, lrs), start=1):
        opt.param_groups[0]["lr"] = lr
        tp.grad = torch.from_numpy(g.copy())
        opt.step()
        before = p.copy()
        p, m, v = fused_math.adam_step_fp32(p, g, m, v, s, lr)
        step_ours, step_torch = p.astype(np.float64) - before, tp.detach().numpy().astype(np.float64) - before
        assert np.allclose(step_ours, step_torch, rtol=1e-5, atol=1e-5 * lr)
        if s == 1:                                              # first step: |step| = lr for ANY non-zero gradient, 1e-12-sized ones too
            assert np.abs(step_torch).min() > 0.99 * lr and np.abs(step_ours).min() > 0.99 * lr


def test_closed_form_sorted_positions_on_random_rectangles():
    """Same property without a renderer: for arbitrary tile rectangles (empty, 1 x N, nested, identical, disjoint) and depths with
    ties, the closed form equals a stable sort of the emitted pairs by (tile id, depth bits)."""
    from tests import mirror_binning as binning
    rng = np.random.default_rng(7)
    gx = 40
    for trial in range(60):
        P = int(rng.integers(1, 21))
        x0 = rng.integers(0, gx - 1, P); y0 = rng.integers(0, 30, P)
        w = rng.integers(0, 7, P); h = rng.integers(0, 7, P)                 # zero width / height: touches no tile
        rects = np.stack([x0, y0, np.minimum(x0 + w, gx), y0 + h], 1)
        rects[(w == 0) | (h == 0)] = 0
        if trial % 3 == 0 and P > 2:
            rects[1] = rects[0]                                               # identical footprints
        depth = rng.integers(0, 4, P).astype(np.uint32) * 1000 + 0x3F800000   # many equal depths
        gs, xy, pos = binning.sorted_positions(rects, depth)
        keys = (xy[:, 1] * gx + xy[:, 0]).astype(np.int64) * (1 << 32) + depth[gs].astype(np.int64)
        expect = np.argsort(keys, kind="stable")                              # emission order breaks ties, as the radix sort does
        assert np.array_equal(np.argsort(pos), expect)


def test_sequence_shards_share_the_rig_and_differ_in_frames():
    """bench.py / multi-GPU runs: rank r optimises shard r of ONE sequence -- same cameras (the scaling loss of round 1 came from
    giving every rank its own rig), independent frames; shard=None reproduces the plain sequence the reference arm optimises."""
    cfg = configs.H36M
    base = synthetic.make_sequence(cfg, 3, seed=100)
    again = synthetic.make_sequence(cfg, 5, seed=100)
    s0, s1 = synthetic.make_sequence(cfg, 3, seed=100, shard=0), synthetic.make_sequence(cfg, 3, seed=100, shard=1)
    for a, b in zip(base.cameras, s1.cameras):
        assert np.array_equal(a.world_view_transform, b.world_view_transform) and np.array_equal(a.full_proj_transform, b.full_proj_transform)
    assert all(np.array_equal(x.pose_3d_gt, y.pose_3d_gt) for x, y in zip(base.frames, again.frames))      # a prefix is a prefix
    assert not np.array_equal(s0.frames[0].pose_3d_gt, s1.frames[0].pose_3d_gt)
    assert not np.array_equal(s0.frames[0].pose_3d_gt, base.frames[0].pose_3d_gt)


def test_adam_step_table_is_torchs_host_arithmetic():
    """training.adam_step_table: the per-step scalars torch/optim/adam.py forms with python floats -- (lr / bc1) * -1 and bc2 ** 0.5
    with bc = 1 - beta ** step -- rounded to fp32; the xyz learning rate is the one set at the stepping iteration."""
    from skelsplat_b200 import training
    cfg = configs.PANOPTIC
    ext = 3217.25
    tab = training.adam_step_table(cfg, ext)
    lr = trainer.xyz_lr_table(cfg, ext)
    assert tab.shape == (125, 5) and tab.dtype == np.float32
    for s in (0, 1, 7, 124):
        step = s + 1
        bc1, bc2 = 1 - 0.9 ** step, 1 - 0.999 ** step
        want = [np.float32((float(lr[step * 4]) / bc1) * -1), np.float32((cfg.scaling_lr / bc1) * -1), np.float32((cfg.rotation_lr / bc1) * -1),
                np.float32((cfg.opacity_lr / bc1) * -1), np.float32(bc2 ** 0.5)]
        assert all(a == b for a, b in zip(tab[s], want))
    assert np.float32(1 - 0.999) != np.float32(1.0) - np.float32(0.999)      # why 1 - beta is formed in fp64 before rounding (csrc/adam_form.h)


def test_factored_heatmap_patches():
    """HeatmapROIs stores a patch as col[h] | row[w]; the heatmap value is their single fp32 product, which is what rois_to_dense
    materialises, so the dense tensors (drop-in path, reference arm) and the fused kernel see identical bits."""
    cfg = configs.OCCLUSION_PERSON
    seq = synthetic.make_sequence(cfg, 1, seed=4)
    fr = seq.frames[0]
    _, scal, rot, _ = trainer.initial_raw_state(cfg, fr.pose_3d_init[None])
    rois = heatmaps.generate_heatmap_rois(fr.pose_3d_init, fr.poses_2d, seq.cameras, scal[0], rot[0])
    assert rois.data.dtype == np.float32 and rois.data.size == int((rois.rect[..., 2] + rois.rect[..., 3]).sum())
    dense = heatmaps.rois_to_dense(rois, 1)
    for j in (0, 7, 14):
        x0, y0, w, h = rois.rect[1, j]
        col, row = rois.factors(1, j)
        assert col.shape == (h,) and row.shape == (w,) and (col > 0).all() and (row > 0).all()
        want = (col[:, None].astype(np.float32) * row[None, :].astype(np.float32)).astype(np.float32)
        assert np.array_equal(dense[j, y0:y0 + h, x0:x0 + w], want)
        assert abs(float(want.max()) - 1.0) < 1e-6
        dense[j, y0:y0 + h, x0:x0 + w] = 0
    assert np.array_equal(heatmaps.heatmap_roi_rects(fr.pose_3d_init, fr.poses_2d, seq.cameras, scal[0], rot[0]), rois.rect)


def test_ssim_epilogue_needs_only_the_sum_of_the_variances():
    """csrc/ssim.cu filters FOUR moments (mu1, mu2, E[x^2] + E[y^2], E[xy]) instead of the reference's five and factors the
    reference's seven divisions into two reciprocals.  Restated in float64: the SSIM value and the three derivative maps of
    submodules/fused-ssim/ssim.cu:262-283 are functions of sigma1^2 + sigma2^2 only, and the factored forms are the same numbers."""
    rng = np.random.default_rng(7)
    n = 10000
    mu1, mu2 = rng.uniform(0, 1, n), rng.uniform(0, 1, n)
    s1, s2 = rng.uniform(0, 0.1, n), rng.uniform(0, 0.1, n)                    # the two variances
    s12 = rng.uniform(-1, 1, n) * np.sqrt(s1 * s2)
    e11, e22, e12 = s1 + mu1 ** 2, s2 + mu2 ** 2, s12 + mu1 * mu2              # the filtered second moments
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    # the reference's expressions, term by term
    Cc, D = 2 * mu1 * mu2 + C1, 2 * (e12 - mu1 * mu2) + C2
    A, B = mu1 ** 2 + mu2 ** 2 + C1, (e11 - mu1 ** 2) + (e22 - mu2 ** 2) + C2
    m_ref = Cc * D / (A * B)
    dmu1_ref = mu2 * 2 * D / (A * B) - mu2 * 2 * Cc / (A * B) - mu1 * 2 * Cc * D / (A * A * B) + mu1 * 2 * Cc * D / (A * B * B)
    ds1_ref, ds12_ref = -Cc * D / (A * B * B), 2 * Cc / (A * B)
    # the kernel's expressions: one summed second moment, two reciprocals
    ess = e11 + e22
    musq = mu1 ** 2 + mu2 ** 2
    A2, B2 = musq + C1, (ess - musq) + C2
    rA, rB = 1 / A2, 1 / B2
    rAB = rA * rB
    m = Cc * D * rAB
    dmu1 = 2 * rAB * (mu2 * (D - Cc) + mu1 * (Cc * D) * (rB - rA))
    ds1, ds12 = -m * rB, 2 * Cc * rAB
    for mine, ref in ((m, m_ref), (dmu1, dmu1_ref), (ds1, ds1_ref), (ds12, ds12_ref)):
        assert np.abs(mine - ref).max() <= 1e-9 * np.abs(ref).max()
    # swapping the two variances (same sum) changes nothing: they enter only as their sum
    B_swapped = (e11 - s1 + s2 - mu1 ** 2) + (e22 - s2 + s1 - mu2 ** 2) + C2
    assert np.allclose(B_swapped, B, rtol=1e-13, atol=0)
