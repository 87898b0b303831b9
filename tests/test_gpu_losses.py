"""Fused dense losses / limb consistency / fused SSIM (through the C ABI) against the restated reference losses
(oracle/pipeline.py = utils/loss_utils.py).  Tolerances: 1e-5 relative; SSIM uses the reference test's own
criterion, torch.isclose with its defaults (submodules/fused-ssim/tests/test.py:82,90)."""
import numpy as np
import pytest
import torch

from oracle import pipeline as opipe
from skelsplat_b200 import configs
from skelsplat_b200 import loss_utils as LU
from tests.util import relerr

pytestmark = pytest.mark.gpu
DEV = "cuda"


def sparse_pair(shape, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    r = torch.rand(shape, generator=g) * (torch.rand(shape, generator=g) > 0.7)
    t = torch.rand(shape, generator=g) * (torch.rand(shape, generator=g) > 0.8)
    return r.to(DEV), t.to(DEV)


@pytest.mark.parametrize("shape", [(17, 250, 333), (19, 31, 7), (15, 64, 64), (1, 1, 1)])
def test_losses_value_and_gradient(shape):
    r, g = sparse_pair(shape, 0)
    cases = [("l2_gaussian", lambda a: LU.l2_loss_gaussian(a, g, None)[0], lambda a: opipe.l2_loss_gaussian(a, g)[0]),
             ("l1", lambda a: LU.l1_loss(a, g, None), lambda a: opipe.l1_loss(a, g)),
             ("l1_gaussian", lambda a: LU.l1_loss_gaussian(a, g, None), lambda a: opipe.l1_loss_gaussian(a, g)),
             ("l1_masked", lambda a: LU.l1_loss_masked(a, g, None), lambda a: opipe.l1_loss_masked(a, g)),
             ("blend", lambda a: LU.l2_loss_gaussian_l1_loss_gaussian(a, g, None, 0.05), lambda a: opipe.l2_loss_gaussian_l1_loss_gaussian(a, g, 0.05)),
             ("l2_gaussian_sum", lambda a: LU.l2_loss_gaussian(a, g, None, reduction="sum"), lambda a: opipe.l2_loss_gaussian(a, g, reduction="sum")),
             ("l1_sum", lambda a: LU.l1_loss(a, g, None, reduction="sum"), lambda a: opipe.l1_loss(a, g, reduction="sum"))]
    for name, mine, ref in cases:
        a = r.clone().requires_grad_(True); b = r.clone().requires_grad_(True)
        lm, lr = mine(a), ref(b)
        (lm * 3.0).backward(); (lr * 3.0).backward()
        assert relerr(lm.item(), lr.item()) < 1e-5, name
        assert relerr(a.grad.cpu().numpy(), b.grad.cpu().numpy()) < 1e-5, name


def test_l2_gaussian_returns_the_error_map_train_py_unpacks():
    r, g = sparse_pair((17, 40, 50), 3)
    loss, err = LU.l2_loss_gaussian(r, g, None, 0.05, reduction="mean")
    assert torch.allclose(err, (r - g) ** 2)
    assert LU.losses["l2_gaussian"] is LU.l2_loss_gaussian and set(LU.consistency_losses) == {"3D_length_consistency", "none"}


def test_empty_mask_gives_nan_like_the_reference():
    z = torch.zeros(3, 8, 8, device=DEV)
    assert torch.isnan(LU.l2_loss_gaussian(z, z, None)[0]) and torch.isnan(opipe.l2_loss_gaussian(z, z)[0])


@pytest.mark.parametrize("name", ["h36m", "panoptic", "occlusion-person"])
def test_limb_consistency(name):
    cfg = configs.get_config(name)
    xyz = torch.randn(6, cfg.n_joints, 3, generator=torch.Generator().manual_seed(1)).to(DEV) * 300
    a = xyz.clone().requires_grad_(True)
    ref = torch.stack([opipe.limb_3d_consistency_loss(a[i], cfg.limb_pairs) for i in range(6)])
    ref.sum().backward()
    b = xyz.clone().requires_grad_(True)
    mine = LU.limb_3d_consistency_loss_batched(b, cfg.limb_pairs)
    mine.sum().backward()
    assert relerr(mine.detach().cpu().numpy(), ref.detach().cpu().numpy()) < 1e-5
    assert relerr(b.grad.cpu().numpy(), a.grad.cpu().numpy()) < 1e-5
    c = xyz[0].clone().requires_grad_(True)
    single = LU.limb_3d_consistency_loss(c, "data/" + cfg.name)          # reference signature (xyz, data_root)
    assert relerr(single.item(), ref[0].item()) < 1e-5


@pytest.mark.parametrize("shape", [(2, 3, 120, 200), (1, 5, 97, 61), (1, 1, 11, 11), (5, 5, 270, 480)])
def test_fused_ssim_against_conv2d_ssim(shape):
    """The reference's own check (tests/test.py:58-91, 'same' padding): value and gradient isclose to conv2d SSIM."""
    from fused_ssim import fused_ssim
    g = torch.Generator().manual_seed(0)
    img1 = torch.rand(shape, generator=g).to(DEV).requires_grad_(True)
    img2 = torch.rand(shape, generator=g).to(DEV)
    a = fused_ssim(img1, img2, "same")
    a.backward()
    ga = img1.grad.clone(); img1.grad = None
    b = opipe.ssim(img1, img2)
    b.backward()
    assert torch.isclose(a, b)
    assert torch.isclose(ga, img1.grad, rtol=1e-5, atol=1e-7).all()
    # 'valid' padding = crop of the same map; inference mode returns the same value without the derivative maps
    if shape[2] > 12:
        v = fused_ssim(img1.detach(), img2, "valid", train=False)
        full = opipe.ssim(img1.detach(), img2, size_average=False)[:, :, 5:-5, 5:-5].mean()
        assert torch.isclose(v, full)
    assert torch.isclose(fused_ssim(img1.detach(), img2, train=False), a.detach())


def test_fused_ssim_identical_images_is_one():
    from fused_ssim import fused_ssim
    x = torch.rand(1, 3, 64, 80, device=DEV)
    assert abs(fused_ssim(x, x.clone()).item() - 1.0) < 1e-6


@pytest.mark.parametrize("shape", [(2, 1, 300, 1501), (1, 3, 65, 33), (3, 2, 64, 256), (1, 1, 129, 31)])
@pytest.mark.parametrize("padding", ["same", "valid"])
def test_fused_ssim_mean_path_equals_the_map_path(shape, padding):
    """fused_ssim() reduces the map inside the kernel (the map and dL/dmap never reach HBM); FusedSSIMMap -- the reference's
    autograd surface, fused_ssim/__init__.py:8-32 -- still materialises it.  Same value and image gradient (summation order
    aside), incl. strips narrower than a warp, heights that are not multiples of the 64-row strips, and the 'valid' crop."""
    import fused_ssim as FS
    g = torch.Generator().manual_seed(4)
    a = torch.rand(shape, generator=g).to(DEV); b = torch.rand(shape, generator=g).to(DEV)
    x = a.clone().requires_grad_(True); y = a.clone().requires_grad_(True)
    v1 = FS.fused_ssim(x, b, padding)
    v2 = FS.FusedSSIMMap.apply(0.01 ** 2, 0.03 ** 2, y, b, padding, True).mean()
    assert torch.isclose(v1, v2, rtol=2e-6, atol=1e-7)
    (3.0 * v1).backward(); (3.0 * v2).backward()
    assert (x.grad - y.grad).abs().max() <= 1e-6 * y.grad.abs().max() + 1e-12
    with pytest.raises(RuntimeError, match="train=True"):
        FS.fused_ssim(a.clone().requires_grad_(True), b, padding, train=False).backward()


def test_adam_kernel_is_bit_exact_against_torch_adam():
    """ssb_adam_frame_step (and phase E of the fused optimiser, which is the same expression) against torch.optim.Adam's own CUDA
    foreach path with the reference's groups / eps (scene/gaussian_model.py:208-218): 125 steps on random gradients incl. zeros
    and noise-level values (what eps = 1e-15 amplifies), xyz learning rate changing every step -- every parameter BIT-equal."""
    import ctypes as C
    from skelsplat_b200 import lib as L_, configs, training, trainer
    cfg = configs.PANOPTIC               # opacity lr != 0
    J, V, n_steps = cfg.n_joints, 4, 125
    g = torch.Generator().manual_seed(5)
    P = [torch.randn(J, 3, generator=g) * 500, torch.full((J, 3), 3.0), torch.randn(J, 4, generator=g), torch.randn(J, 1, generator=g)]
    ref = [torch.nn.Parameter(p.clone().to(DEV)) for p in P]
    names = ("xyz", "scaling", "rotation", "opacity")
    ext = 3217.25
    opt = torch.optim.Adam([{"params": [ref[0]], "lr": cfg.position_lr_init * ext, "name": "xyz"}, {"params": [ref[3]], "lr": cfg.opacity_lr, "name": "opacity"},
                            {"params": [ref[1]], "lr": cfg.scaling_lr, "name": "scaling"}, {"params": [ref[2]], "lr": cfg.rotation_lr, "name": "rotation"}],
                           lr=0.0, eps=1e-15)
    mine = [p.clone().to(DEV) for p in P]
    table = torch.from_numpy(training.adam_step_table(cfg, ext)).to(DEV)
    lr = trainer.xyz_lr_table(cfg, ext)
    exp_avg = torch.zeros(11 * J, device=DEV); exp_avg_sq = torch.zeros(11 * J, device=DEV); counter = torch.zeros(1, dtype=torch.int32, device=DEV)
    lib = L_.lib()
    for s in range(n_steps):
        acc = torch.randn(V, J, 3, generator=g).to(DEV) * (10.0 ** float(torch.randint(-12, 2, (1,), generator=g)))
        if s % 7 == 0:
            acc[1] = 0
        gs = [torch.randn(J, 3, generator=g).to(DEV) * 1e-3, torch.randn(J, 4, generator=g).to(DEV) * 1e-11, torch.zeros(J, 1, device=DEV)]
        if s % 5 == 0:
            gs[2] = torch.randn(J, 1, generator=g).to(DEV) * 1e-6
        for grp in opt.param_groups:
            if grp["name"] == "xyz":
                grp["lr"] = float(lr[(s + 1) * 4])
        ref[0].grad = acc.mean(dim=0); ref[1].grad, ref[2].grad, ref[3].grad = gs[0].clone(), gs[1].clone(), gs[2].clone()
        opt.step()
        p_ = L_.ptr
        L_.check(lib.ssb_adam_frame_step(C.c_int(J), C.c_int(V), p_(mine[0]), p_(mine[1]), p_(mine[2]), p_(mine[3]), p_(acc), p_(gs[0]), p_(gs[1]), p_(gs[2]),
                                         p_(exp_avg), p_(exp_avg_sq), p_(table), C.c_int(n_steps), p_(counter), C.c_float(float(np.float32(1 - 0.9))),
                                         C.c_float(0.999), C.c_float(float(np.float32(1 - 0.999))), C.c_float(1e-15), L_.current_stream()), "adam")
        for n, a, b in zip(names, mine, ref):
            assert torch.equal(a, b.detach()), (n, s, float((a - b.detach()).abs().max()))
    assert int(counter.item()) == n_steps
