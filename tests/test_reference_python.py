"""The reference's OWN Python (imported unmodified through tests/ref_import.py) against (a) this repo's mirrors and the
drop-in packages -- the boundary proven through the reference's own caller -- and (b) oracle/pipeline.py, which pins the
restated oracle to the files it restates.

CPU tests read /root/reference (or the staged copy) and skip when neither exists; GPU tests use the staged copy
oracle/_ref/pyref, which travels to the GPU box next to the compiled reference kernels."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import pipeline as opipe
from skelsplat_b200 import configs, synthetic, trainer, heatmaps
from skelsplat_b200.cameras import cameras_extent
from tests import ref_import
from tests.util import small_config

needs_ref = pytest.mark.skipif(not ref_import.available(), reason="reference Python not present (no /root/reference, no oracle/_ref/pyref)")
DEV = "cuda"


# ------------------------------------------------------------------------------------------------ CPU: oracle pinning
@needs_ref
def test_oracle_losses_equal_the_reference_functions():
    """oracle/pipeline.py's restated losses are the reference's (utils/loss_utils.py:67-127,173-192,226-250): bit-equal values AND
    gradients on random inputs with empty / partial / full masks."""
    ref = ref_import.load("ours")
    LU = ref.loss_utils
    g = torch.Generator().manual_seed(0)
    for shape, sparsity in (((5, 40, 37), 0.8), ((3, 16, 16), 0.0), ((2, 9, 11), 0.97)):
        gt = torch.rand(shape, generator=g) * (torch.rand(shape, generator=g) > sparsity)
        base = torch.rand(shape, generator=g) * (torch.rand(shape, generator=g) > sparsity)
        for ref_fn, my_fn, tuple_out in ((LU.l2_loss_gaussian, opipe.l2_loss_gaussian, True), (LU.l1_loss_gaussian, opipe.l1_loss_gaussian, False),
                                         (LU.l1_loss, opipe.l1_loss, False), (LU.l1_loss_masked, opipe.l1_loss_masked, False)):
            a = base.clone().requires_grad_(True); b = base.clone().requires_grad_(True)
            ra = ref_fn(a, gt, None, 0.05, reduction="mean"); rb = my_fn(b, gt, reduction="mean")
            if tuple_out:
                assert torch.equal(ra[1], rb[1])
                ra, rb = ra[0], rb[0]
            assert torch.equal(ra, rb)
            ra.backward(); rb.backward()
            assert torch.equal(a.grad, b.grad)
        for lam in (0.0, 0.05, 1.0):
            assert torch.equal(LU.l2_loss_gaussian_l1_loss_gaussian(base, gt, None, lam), opipe.l2_loss_gaussian_l1_loss_gaussian(base, gt, lam))
    for cfg in (configs.H36M, configs.PANOPTIC, configs.OCCLUSION_PERSON):
        x = (torch.randn(cfg.n_joints, 3, generator=g) * 300).requires_grad_(True)
        y = x.detach().clone().requires_grad_(True)
        ra = LU.limb_3d_consistency_loss(x, "data/" + cfg.name); rb = opipe.limb_3d_consistency_loss(y, cfg.limb_pairs)
        assert torch.equal(ra, rb)
        ra.backward(); rb.backward()
        assert torch.equal(x.grad, y.grad)


@needs_ref
def test_lr_schedule_ssim_and_registries_equal_the_reference():
    ref = ref_import.load("ours")
    GU, LU = ref.general_utils, ref.loss_utils
    for cfg, ext in ((configs.H36M, 5234.5), (configs.PANOPTIC, 3100.25)):
        f = GU.get_expon_lr_func(lr_init=cfg.position_lr_init * ext, lr_final=cfg.position_lr_final * ext,
                                 lr_delay_mult=cfg.position_lr_delay_mult, max_steps=cfg.position_lr_max_steps)
        o = opipe.get_expon_lr_func(cfg.position_lr_init * ext, cfg.position_lr_final * ext, lr_delay_mult=cfg.position_lr_delay_mult,
                                    max_steps=cfg.position_lr_max_steps)
        tab = trainer.xyz_lr_table(cfg, ext)
        assert all(f(i) == o(i) == tab[i] for i in range(1, 501))         # fp64 host scalars, identical expression
    assert torch.isinf(GU.inverse_sigmoid(torch.ones(1))).all()
    g = torch.Generator().manual_seed(1)
    a, b = torch.rand(2, 3, 40, 52, generator=g), torch.rand(2, 3, 40, 52, generator=g)
    assert torch.equal(LU.ssim(a, b), opipe.ssim(a, b))                    # the conv2d SSIM fused-ssim's own test is pinned to
    # train.py:150 unpacks a 2-tuple: l2_gaussian is the only entry of the reference's table that returns one
    x, y = torch.rand(3, 8, 8, generator=g), torch.rand(3, 8, 8, generator=g)
    for name, fn in ref.utils.losses.items():
        if name in ("l2", "l2_sqrt", "huber", "l1_l2", "l1_huber", "l1_masked_l2", "l1_masked_huber", "cauchy"):
            continue                                                       # need 2D keypoints / softargmax (never configured)
        out = fn(x, y, None, 0.05, reduction="mean")
        assert isinstance(out, tuple) == (name == "l2_gaussian")
    assert set(ref.utils.consistency_losses) == {"3D_length_consistency", "none"}


@needs_ref
def test_mirror_tables_cover_the_reference_registries():
    """The mirror's registries carry the keys the shipped configs select, under the reference's names."""
    ref = ref_import.load("ours")
    from skelsplat_b200 import loss_utils as mine
    for cfg in configs.CONFIGS.values():
        assert cfg.loss_function in ref.utils.losses and cfg.loss_function in mine.losses
        assert cfg.consistency_loss in ref.utils.consistency_losses and cfg.consistency_loss in mine.consistency_losses
        assert cfg.rendering in ref.gaussian_renderer.render_functions


# ------------------------------------------------------------------------------------------------ GPU
def _ref_model(ref, cfg, frame, cams, iterations=500):
    gm = ref.gaussian_model.GaussianModel(1, "default")
    pts = np.asarray(frame.pose_3d_init, np.float32)
    pcd = SimpleNamespace(points=pts, colors=np.zeros_like(pts), normals=np.zeros_like(pts))
    gm.create_from_pcd(pcd, [SimpleNamespace(image_name=f"c{c.uid}") for c in cams], cameras_extent(cams), cfg.opacity_on, cfg.scaling,
                       cfg.n_joints, cfg.scaling_modifier, cfg.name)
    opt = SimpleNamespace(position_lr_init=cfg.position_lr_init, position_lr_final=cfg.position_lr_final,
                          position_lr_delay_mult=cfg.position_lr_delay_mult, position_lr_max_steps=cfg.position_lr_max_steps,
                          feature_lr=cfg.feature_lr, opacity_lr=cfg.opacity_lr, scaling_lr=cfg.scaling_lr, rotation_lr=cfg.rotation_lr,
                          percent_dense=0.01, exposure_lr_init=0.01, exposure_lr_final=0.001, exposure_lr_delay_steps=0,
                          exposure_lr_delay_mult=0.0, iterations=iterations)
    gm.training_setup(opt)
    return gm


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("name", ["h36m", "h36m-occ", "panoptic", "occlusion-person"])
def test_reference_gaussian_model_equals_mirror_and_oracle(name):
    """scene/gaussian_model.py:149-248 (create_from_pcd, training_setup, update_learning_rate) run unmodified: its tensors, Adam
    groups and learning rates equal the mirror's (skelsplat_b200/gaussian_model.py), the oracle's (RefGaussianModel) and the
    packed initial state of the fused path (trainer.initial_raw_state)."""
    from skelsplat_b200.gaussian_model import GaussianModel as Mirror
    ref = ref_import.load("ours")
    cfg = configs.get_config(name)
    seq = synthetic.make_sequence(cfg, 1, seed=5)
    fr, cams = seq.frames[0], seq.cameras
    gm = _ref_model(ref, cfg, fr, cams)
    mi = Mirror(1, "default", DEV)
    mi.create_from_pcd(np.asarray(fr.pose_3d_init, np.float32), cams, cameras_extent(cams), cfg.opacity_on, cfg.scaling, cfg.n_joints,
                       cfg.scaling_modifier, cfg.name)
    mi.training_setup(SimpleNamespace(position_lr_init=cfg.position_lr_init, position_lr_final=cfg.position_lr_final,
                                      position_lr_delay_mult=cfg.position_lr_delay_mult, position_lr_max_steps=cfg.position_lr_max_steps,
                                      feature_lr=cfg.feature_lr, opacity_lr=cfg.opacity_lr, scaling_lr=cfg.scaling_lr, rotation_lr=cfg.rotation_lr,
                                      percent_dense=0.01))
    orc = opipe.RefGaussianModel(fr.pose_3d_init, cfg, cameras_extent(cams), DEV)
    xyz0, scal0, rot0, opa0 = trainer.initial_raw_state(cfg, fr.pose_3d_init[None])
    for attr, packed in (("_xyz", xyz0[0]), ("_scaling", scal0[0]), ("_rotation", rot0[0]), ("_opacity", opa0[0][:, None])):
        a = getattr(gm, attr).detach()
        assert torch.equal(a, getattr(mi, attr).detach()) and torch.equal(a, getattr(orc, attr).detach()), attr
        assert np.array_equal(a.cpu().numpy(), packed), attr
    for getter in ("get_xyz", "get_scaling", "get_rotation", "get_opacity", "get_features"):
        a = getattr(gm, getter).detach()
        assert torch.equal(a, getattr(mi, getter).detach()) and torch.equal(a, getattr(orc, getter).detach()), getter
    groups = lambda o: [(g["name"], g["lr"], tuple(g["params"][0].shape)) for g in o.optimizer.param_groups]
    assert groups(gm) == groups(mi) == groups(orc)
    assert gm.optimizer.defaults["eps"] == mi.optimizer.defaults["eps"] == 1e-15
    tab = trainer.xyz_lr_table(cfg, cameras_extent(cams))
    for it in (1, 4, 8, 250, 500):
        assert gm.update_learning_rate(it) == mi.update_learning_rate(it) == orc.update_learning_rate(it) == tab[it]


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("name", ["h36m", "panoptic", "occlusion-person"])
def test_reference_render_and_losses_on_the_dropin_packages(name):
    """gaussian_renderer.render_* (gaussian_renderer/__init__.py:28-364) + utils.loss_utils run UNMODIFIED on this repo's
    diff_gaussian_rasterization_* packages: outputs and gradients are bit-equal to the mirror's render_* (same op underneath;
    the mirror's fused loss kernel within 1e-6), and equal -- keys / radii exactly, image <= 1e-5 -- to the same reference
    Python on the reference's own kernels."""
    from oracle import ref_rasterizer as refr
    from skelsplat_b200.gaussian_renderer import render_functions as mirror_render
    from skelsplat_b200 import loss_utils as mirror_losses
    from skelsplat_b200.training import TorchCamera
    cfg = configs.get_config(name)
    ours = ref_import.load("ours")
    seq = synthetic.make_sequence(cfg, 1, seed=6)
    fr, cams = seq.frames[0], seq.cameras
    pipe = SimpleNamespace(debug=False, antialiasing=False, compute_cov3D_python=False, convert_SHs_python=False)
    bg = torch.tensor([0, 0, 0], dtype=torch.float32, device=DEV)
    _, scal0, rot0, _ = trainer.initial_raw_state(cfg, fr.pose_3d_init[None])
    rois = heatmaps.generate_heatmap_rois(fr.pose_3d_init, fr.poses_2d, cams, scal0[0], rot0[0])
    variant = opipe.VARIANT_OF[cfg.rendering]
    for v in (0, len(cams) - 1):
        cam = TorchCamera(cams[v], DEV)
        gt = torch.from_numpy(heatmaps.rois_to_dense(rois, v)).to(DEV)
        outs = {}
        for tag, render, loss_fn, cons_fn in (
                ("ref-python/ours", ours.gaussian_renderer.render_functions[cfg.rendering], ours.utils.losses["l2_gaussian"],
                 ours.utils.consistency_losses["3D_length_consistency"]),
                ("mirror/ours", mirror_render[cfg.rendering], mirror_losses.losses["l2_gaussian"], mirror_losses.consistency_losses["3D_length_consistency"])):
            gm = _ref_model(ours, cfg, fr, cams)
            with torch.no_grad():       # leave the symmetric initial state: every gradient becomes non-trivial
                gm._scaling.add_(torch.linspace(-0.3, 0.3, gm._scaling.numel(), device=DEV).reshape(gm._scaling.shape))
                gm._rotation.add_(torch.linspace(-0.2, 0.2, gm._rotation.numel(), device=DEV).reshape(gm._rotation.shape))
            pkg = render(cam, gm, pipe, bg)
            l2, err = loss_fn(pkg["render"], gt, None, cfg.lambda_loss_function, reduction="mean")
            loss = l2 + cons_fn(gm.get_xyz, "data/" + cfg.name, reduction="mean") * cfg.lambda_consistency
            grads = torch.autograd.grad(loss, [gm.get_xyz, gm._scaling, gm._rotation])
            outs[tag] = (pkg, l2.detach(), [g.detach() for g in grads])
        a, b = outs["ref-python/ours"], outs["mirror/ours"]
        assert set(a[0].keys()) == set(b[0].keys()) == {"render", "viewspace_points", "visibility_filter", "radii", "depth"}
        assert torch.equal(a[0]["render"], b[0]["render"]) and torch.equal(a[0]["radii"], b[0]["radii"]) and torch.equal(a[0]["depth"], b[0]["depth"])
        assert torch.equal(a[0]["visibility_filter"], b[0]["visibility_filter"])
        assert abs(float(a[1]) - float(b[1])) <= 1e-6 * abs(float(a[1]))
        for ga, gb in zip(a[2], b[2]):
            assert (ga - gb).abs().max() <= 1e-5 * ga.abs().max() + 1e-30
        if refr.available(variant):
            theirs = ref_import.load("ref")
            runs = []
            for _ in range(2):           # two runs: the reference's backward uses unordered fp32 atomics -- its own spread sets the gradient tolerance
                gm = _ref_model(theirs, cfg, fr, cams)
                with torch.no_grad():
                    gm._scaling.add_(torch.linspace(-0.3, 0.3, gm._scaling.numel(), device=DEV).reshape(gm._scaling.shape))
                    gm._rotation.add_(torch.linspace(-0.2, 0.2, gm._rotation.numel(), device=DEV).reshape(gm._rotation.shape))
                pkg = theirs.gaussian_renderer.render_functions[cfg.rendering](cam, gm, pipe, bg)
                l2, _ = theirs.utils.losses["l2_gaussian"](pkg["render"], gt, None, cfg.lambda_loss_function, reduction="mean")
                loss = l2 + theirs.utils.consistency_losses["3D_length_consistency"](gm.get_xyz, "data/" + cfg.name, reduction="mean") * cfg.lambda_consistency
                runs.append([g.detach() for g in torch.autograd.grad(loss, [gm.get_xyz, gm._scaling, gm._rotation])])
            grads = runs[0]
            assert torch.equal(pkg["radii"], a[0]["radii"])
            ref_img = pkg["render"]
            assert (ref_img - a[0]["render"]).abs().max() <= 1e-5 * ref_img.abs().max()
            assert torch.equal(ref_img > 0, a[0]["render"] > 0)             # the loss mask is the same set of pixels
            assert abs(float(l2) - float(a[1])) <= 1e-5 * abs(float(l2))
            for gr, gr2, ga in zip(grads, runs[1], a[2]):
                spread = float((gr - gr2).abs().max() / gr.abs().max())
                assert float((gr - ga).abs().max() / gr.abs().max()) <= max(2e-5, 4.0 * spread), (name, v, spread)


@pytest.mark.gpu
@needs_ref
def test_reference_generate_heatmaps_equals_the_roi_specification():
    """utils/general_utils.py:175-304 run unmodified (scipy's gaussian_filter standing in for cupyx's) against heatmaps.py:
    the reference evaluates sigma in fp32 on the GPU, the specification in fp64 with a fixed operation order, so a window may
    differ by 1 px where 4 sigma + 0.5 sits on an integer (rare); everywhere else the maps agree to 2e-6 and the masks exactly."""
    ref = ref_import.load("ours")
    cfg = configs.H36M
    seq = synthetic.make_sequence(cfg, 2, seed=8)
    from skelsplat_b200.training import TorchCamera
    n_same, n_tot = 0, 0
    for fr in seq.frames:
        gm = _ref_model(ref, cfg, fr, seq.cameras)
        tcams = [TorchCamera(c, DEV) for c in seq.cameras]
        cov = ref.general_utils.unpack_covariance(gm.get_covariance())                  # train.py:91-92
        p2d = torch.as_tensor(np.asarray(fr.poses_2d, np.float32)).to(DEV)
        hm = ref.general_utils.generate_heatmaps(gm, p2d, tcams, cov, False, "data/" + cfg.name, cfg.nviews)
        _, scal0, rot0, _ = trainer.initial_raw_state(cfg, fr.pose_3d_init[None])
        rois = heatmaps.generate_heatmap_rois(fr.pose_3d_init, fr.poses_2d, seq.cameras, scal0[0], rot0[0])
        for v in range(cfg.nviews):
            theirs = hm[str(v)].cpu().numpy(); mine = heatmaps.rois_to_dense(rois, v)
            assert theirs.shape == mine.shape
            for j in range(cfg.n_joints):
                n_tot += 1
                if np.array_equal(theirs[j] > 0, mine[j] > 0):
                    n_same += 1
                    assert np.abs(theirs[j] - mine[j]).max() < 2e-6
                else:       # a 1-px window difference: the interiors still agree
                    both = (theirs[j] > 0) & (mine[j] > 0)
                    assert np.abs(theirs[j] - mine[j])[both].max() < 1e-3 and abs(int((theirs[j] > 0).sum()) - int((mine[j] > 0).sum())) < 200
    assert n_same >= 0.9 * n_tot


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("name", ["h36m", "occlusion-person"])
def test_reference_python_loop_on_the_dropin_equals_fused_optimiser(name):
    """train.py's iteration body (skelsplat_b200/training.py restates it; the loop itself cannot be imported: hydra) driving the
    REFERENCE's GaussianModel / render_* / l2_loss_gaussian / limb_3d_consistency_loss on the drop-in packages, against the
    same loop on the mirrors and against the fused kernel."""
    from skelsplat_b200.training import optimise_frame_dropin
    ref = ref_import.load("ours")
    cfg = small_config(configs.get_config(name), factor=2)
    seq = synthetic.make_sequence(cfg, 1, seed=9)
    fr = seq.frames[0]
    mods = (ref.gaussian_model.GaussianModel, ref.gaussian_renderer.render_functions, ref.utils.losses, ref.utils.consistency_losses)
    iters = 40
    a = optimise_frame_dropin(fr, seq.cameras, cfg, device=DEV, iterations=iters, modules=mods)
    b = optimise_frame_dropin(fr, seq.cameras, cfg, device=DEV, iterations=iters)
    c = trainer.optimize_sequence(seq, DEV, iterations=iters)[0]
    assert np.linalg.norm(a - fr.pose_3d_init, axis=-1).max() > 1.0
    assert np.linalg.norm(a - b, axis=-1).max() < 0.02          # reference Python vs mirrors: same op, torch loss vs fused loss kernel
    assert np.linalg.norm(a - c, axis=-1).max() < 0.05          # vs the fused optimiser


@pytest.mark.gpu
@needs_ref
def test_oracle_loop_equals_the_reference_python_on_the_reference_kernels():
    """Pins oracle/pipeline.py's restated loop: the reference's own Python on the reference's own kernels (backend "ref")
    produces the same poses as the restated loop on the same kernels, up to the kernels' own atomics noise (8 iterations keep
    the Adam amplification of that noise small)."""
    from oracle import ref_rasterizer as refr
    from skelsplat_b200.training import optimise_frame_dropin
    if not refr.available("h36m"):
        pytest.skip("oracle/_ref not present on this box")
    ref = ref_import.load("ref")
    cfg = configs.H36M
    seq = synthetic.make_sequence(cfg, 1, seed=10)
    fr = seq.frames[0]
    _, scal0, rot0, _ = trainer.initial_raw_state(cfg, fr.pose_3d_init[None])
    rois = heatmaps.generate_heatmap_rois(fr.pose_3d_init, fr.poses_2d, seq.cameras, scal0[0], rot0[0])
    dense = [torch.from_numpy(heatmaps.rois_to_dense(rois, v)).to(DEV) for v in range(cfg.nviews)]
    mods = (ref.gaussian_model.GaussianModel, ref.gaussian_renderer.render_functions, ref.utils.losses, ref.utils.consistency_losses)
    a = optimise_frame_dropin(fr, seq.cameras, cfg, heatmaps_dense=dense, device=DEV, iterations=8, modules=mods)
    b = opipe.optimise_frame(fr, seq.cameras, cfg, cameras_extent(seq.cameras), dense, backend="ref", device=DEV, iterations=8)
    assert np.linalg.norm(b - fr.pose_3d_init, axis=-1).max() > 1.0
    assert np.linalg.norm(a - b, axis=-1).max() < 0.02


@pytest.mark.gpu
@pytest.mark.skipif(not ref_import.ref_ssim_available(), reason="oracle/_ref/fused_ssim_cuda.so (make -C oracle ref_ssim) not present")
@pytest.mark.parametrize("shape,padding", [((2, 3, 97, 130), "same"), ((1, 1, 64, 64), "same"), ((5, 1, 300, 301), "valid"), ((1, 17, 40, 33), "valid")])
def test_fused_ssim_against_the_reference_kernels(shape, padding):
    """csrc/ssim.cu against fused-ssim's OWN kernels (submodules/fused-ssim/ssim.cu:187-444 compiled unmodified) under the
    reference's own fused_ssim/__init__.py: value, map-level derivative tensors and the image gradient."""
    import fused_ssim as mine
    theirs = ref_import.load_fused_ssim("ref")
    g = torch.Generator(device="cpu").manual_seed(3)
    a = torch.rand(shape, generator=g).to(DEV); b = torch.rand(shape, generator=g).to(DEV)
    x = a.clone().requires_grad_(True); y = a.clone().requires_grad_(True)
    vr = theirs.fused_ssim(x, b, padding=padding); vm = mine.fused_ssim(y, b, padding=padding)
    assert torch.isclose(vr, vm, rtol=1e-5, atol=1e-7)
    vr.backward(); vm.backward()
    assert (x.grad - y.grad).abs().max() <= 1e-5 * x.grad.abs().max() + 1e-9
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    mr = theirs.fusedssim(C1, C2, a, b, True); mm = mine.fusedssim(C1, C2, a, b, True)
    for tr, tm in zip(mr, mm):
        assert (tr - tm).abs().max() <= 2e-5 * max(1.0, float(tr.abs().max()))
    # inference mode returns the same map
    assert torch.equal(mine.fusedssim(C1, C2, a, b, False)[0], mm[0])
    # the reference wrapper runs unchanged on OUR extension surface (drop-in at the fused_ssim_cuda level)
    swapped = ref_import.load_fused_ssim("ours")
    z = a.clone().requires_grad_(True)
    vs = swapped.fused_ssim(z, b, padding=padding)          # their wrapper: map.mean() in torch; ours reduces inside the kernel
    assert torch.isclose(vs, vm, rtol=2e-6, atol=1e-7)
    vs.backward()
    assert (z.grad - y.grad).abs().max() <= 1e-6 * y.grad.abs().max() + 1e-12
