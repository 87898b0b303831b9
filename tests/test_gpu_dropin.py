"""The drop-in surface: reference-named packages, error behaviour, render_* dict, and one full
train.py-style iteration (render -> l2_gaussian + consistency -> autograd) against the oracle loop."""
import numpy as np
import pytest
import torch

from oracle import pipeline as opipe
from skelsplat_b200 import configs, synthetic, heatmaps, trainer
from skelsplat_b200.cameras import cameras_extent
from tests.util import small_config, relerr

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_packages_and_channel_counts():
    import diff_gaussian_rasterization_h36m as a, diff_gaussian_rasterization_panoptic as b, diff_gaussian_rasterization_op as c
    assert (a.NUM_CHANNELS, b.NUM_CHANNELS, c.NUM_CHANNELS) == (17, 19, 15)
    fields = ("image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix", "projmatrix", "sh_degree",
              "campos", "prefiltered", "debug", "antialiasing")
    for m in (a, b, c):
        assert m.GaussianRasterizationSettings._fields == fields            # RAST/.../__init__.py:143-156


def _settings(mod, cam, dev=DEV):
    t = lambda x: torch.from_numpy(x).to(dev)
    return mod.GaussianRasterizationSettings(cam.image_height, cam.image_width, cam.tanfovx, cam.tanfovy, torch.zeros(3, device=dev), 1.0,
                                             t(cam.world_view_transform), t(cam.full_proj_transform), 0, t(cam.camera_center), False, True, False)


def test_argument_validation_matches_the_reference():
    import diff_gaussian_rasterization_h36m as m
    cfg = small_config(configs.H36M)
    seq = synthetic.make_sequence(cfg, 1, seed=0)
    rast = m.GaussianRasterizer(_settings(m, seq.cameras[0]))
    J = 17
    x = torch.zeros(J, 3, device=DEV); o = torch.ones(J, 1, device=DEV); f = torch.eye(J, device=DEV).reshape(J, 1, J)
    s = torch.ones(J, 3, device=DEV); q = torch.zeros(J, 4, device=DEV); q[:, 0] = 1
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        rast(x, x, o, shs=None, colors_precomp=None, scales=s, rotations=q)
    with pytest.raises(Exception, match="excatly one"):
        rast(x, x, o, shs=f, colors_precomp=f, scales=s, rotations=q)
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        rast(x, x, o, shs=f, scales=s, rotations=None)
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        rast(x, x, o, shs=f, scales=s, rotations=q, cov3D_precomp=torch.zeros(J, 6, device=DEV))
    with pytest.raises(RuntimeError, match="means3D must have dimensions"):
        rast(torch.zeros(J, 4, device=DEV), x, o, shs=f, scales=s, rotations=q)
    with pytest.raises(RuntimeError, match="17 channels"):
        rast(x, x, o, shs=torch.zeros(J, 1, 3, device=DEV), scales=s, rotations=q)
    # P == 0: nothing is launched, zero images come back (RAST/rasterize_points.cu:88)
    e = torch.zeros(0, 3, device=DEV)
    color, radii, invd = rast(e, e, torch.zeros(0, 1, device=DEV), shs=torch.zeros(0, 1, J, device=DEV), scales=e, rotations=torch.zeros(0, 4, device=DEV))
    assert color.shape == (J, seq.cameras[0].image_height, seq.cameras[0].image_width) and not color.any() and radii.numel() == 0


@pytest.mark.parametrize("name", ["h36m", "panoptic", "occlusion-person"])
def test_one_training_iteration_matches_the_oracle_loop(name):
    """render_* -> clamp -> l2_gaussian + 1e-5*consistency -> autograd.grad, vs the restated loop on the C oracle."""
    from skelsplat_b200.gaussian_model import GaussianModel
    from skelsplat_b200.gaussian_renderer import render_functions
    from skelsplat_b200.loss_utils import losses, consistency_losses
    from skelsplat_b200.training import TorchCamera
    from types import SimpleNamespace
    cfg = small_config(configs.get_config(name))
    seq = synthetic.make_sequence(cfg, 1, seed=5)
    fr = seq.frames[0]
    ext = cameras_extent(seq.cameras)
    _, scal0, rot0, _ = trainer.initial_raw_state(cfg, fr.pose_3d_init[None])
    rois = heatmaps.generate_heatmap_rois(fr.pose_3d_init, fr.poses_2d, seq.cameras, scal0[0], rot0[0])
    pipe = SimpleNamespace(debug=True, antialiasing=False, compute_cov3D_python=False, convert_SHs_python=False)
    bg = torch.zeros(3, device=DEV)
    for v in (0, 2):
        gm = GaussianModel(1, "default", DEV)
        gm.create_from_pcd(fr.pose_3d_init.astype(np.float32), seq.cameras, ext, True, cfg.scaling, cfg.n_joints, cfg.scaling_modifier, cfg.name)
        with torch.no_grad():                               # leave the symmetric initial state: exercise scale / rotation gradients
            gm._scaling += torch.linspace(-0.4, 0.9, cfg.n_joints * 3, device=DEV).reshape(-1, 3)
            gm._rotation += 0.3 * torch.sin(torch.arange(cfg.n_joints * 4, device=DEV).float()).reshape(-1, 4)
        gt = torch.from_numpy(heatmaps.rois_to_dense(rois, v)).to(DEV)
        pkg = render_functions[cfg.rendering](TorchCamera(seq.cameras[v], DEV), gm, pipe, bg)
        assert set(pkg.keys()) == {"render", "viewspace_points", "visibility_filter", "radii", "depth"}
        assert pkg["visibility_filter"].shape[1] == 1
        l2, err = losses[cfg.loss_function](pkg["render"], gt, None, cfg.lambda_loss_function, reduction="mean")
        loss = l2 + consistency_losses[cfg.consistency_loss](gm.get_xyz, "data/" + cfg.name, reduction="mean") * cfg.lambda_consistency
        g_mine = torch.autograd.grad(loss, [gm.get_xyz, gm._scaling, gm._rotation, gm._opacity])
        # oracle side (CPU): same parameters
        om = opipe.RefGaussianModel(fr.pose_3d_init, cfg, ext, "cpu")
        with torch.no_grad():
            om._scaling.copy_(gm._scaling.cpu()); om._rotation.copy_(gm._rotation.cpu())
        opkg = opipe.render(opipe.TorchCamera(seq.cameras[v], "cpu"), om, torch.zeros(3), "oracle", opipe.VARIANT_OF[cfg.rendering])
        ol2, _ = opipe.l2_loss_gaussian(opkg["render"], gt.cpu())
        oloss = ol2 + opipe.limb_3d_consistency_loss(om.get_xyz, cfg.limb_pairs) * cfg.lambda_consistency
        g_ref = torch.autograd.grad(oloss, [om.get_xyz, om._scaling, om._rotation, om._opacity])
        assert relerr(pkg["render"].detach().cpu().numpy(), opkg["render"].detach().numpy()) < 1e-5
        assert torch.equal(pkg["radii"].cpu(), opkg["radii"])
        assert relerr(loss.item(), oloss.item()) < 1e-5
        for a, b, n in zip(g_mine, g_ref, ("xyz", "scaling", "rotation", "opacity")):
            assert relerr(a.cpu().numpy(), b.numpy()) < 3e-5, n
        assert not g_mine[3].any()                          # sigmoid'(+inf) == 0: opacity never moves


def test_dropin_loop_equals_fused_optimiser():
    """The per-iteration drop-in loop (dense images, torch Adam) and the fused persistent kernel implement the
    same algorithm: 40 iterations on a non-chaotic config (rotation lr 0) agree to < 0.01 mm."""
    from skelsplat_b200.training import optimise_frame_dropin
    cfg = small_config(configs.OCCLUSION_PERSON)
    seq = synthetic.make_sequence(cfg, 2, seed=6)
    fused = trainer.optimize_sequence(seq, DEV, iterations=40)
    for fi, fr in enumerate(seq.frames):
        drop = optimise_frame_dropin(fr, seq.cameras, cfg, device=DEV, iterations=40)
        assert np.linalg.norm(drop - fused[fi], axis=-1).max() < 0.01


@pytest.mark.parametrize("name", ["h36m", "panoptic", "occlusion-person-8v"])
def test_graphed_frame_optimizer_equals_the_eager_dense_loop(name):
    """training.GraphedFrameOptimizer -- one CUDA graph per Adam step (4 iteration bodies + the Adam kernel), captured once per
    rig and replayed for every frame -- against optimise_frame_dropin (eager launches, torch.optim.Adam): the same poses to
    fp32 rounding of the clamp-free graph (bit-identical in practice), for consecutive frames through the SAME captured graphs
    (static buffers re-initialised in place), heatmaps given densely or as factored ROIs."""
    from skelsplat_b200.training import optimise_frame_dropin, GraphedFrameOptimizer
    cfg = small_config(configs.get_config(name), factor=2)
    seq = synthetic.make_sequence(cfg, 3, seed=15)
    iters = 24
    gfo = GraphedFrameOptimizer(cfg, seq.cameras, DEV, iterations=iters)
    for fi, fr in enumerate(seq.frames):
        _, scal0, rot0, _ = trainer.initial_raw_state(cfg, fr.pose_3d_init[None])
        rois = heatmaps.generate_heatmap_rois(fr.pose_3d_init, fr.poses_2d, seq.cameras, scal0[0], rot0[0])
        dense = [torch.from_numpy(heatmaps.rois_to_dense(rois, v)).to(DEV) for v in range(cfg.nviews)]
        eager = optimise_frame_dropin(fr, seq.cameras, cfg, heatmaps_dense=dense, device=DEV, iterations=iters)
        got = gfo.optimise(fr.pose_3d_init, dense=dense) if fi == 1 else gfo.optimise(fr.pose_3d_init, rois=rois)
        assert np.linalg.norm(eager - fr.pose_3d_init, axis=-1).max() > 1.0
        assert np.linalg.norm(got - eager, axis=-1).max() < 1e-3, (name, fi)
    assert len(gfo.graphs) == (2 if cfg.nviews == 8 else 1)
