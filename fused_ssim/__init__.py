"""Drop-in for the reference's ``fused_ssim`` package (submodules/fused-ssim/fused_ssim/__init__.py:8-41),
backed by skelsplat_b200's sm_100a kernels (csrc/ssim.cu) through the C ABI.

``fused_ssim(img1, img2, padding="same", train=True)`` -> scalar mean SSIM; gradient flows to ``img1``
only, as in the reference (its backward returns ``None`` for img2)."""
import ctypes as C

import torch

from skelsplat_b200 import lib as _L

allowed_padding = ["same", "valid"]


def fusedssim(C1, C2, img1, img2, train=True):
    """fused_ssim_cuda.fusedssim (submodules/fused-ssim/ssim.cu:368-404): returns
    (ssim_map, dm_dmu1, dm_dsigma1_sq, dm_dsigma12); the last three are empty when train is False."""
    L = _L.lib()
    img1 = img1.contiguous().float(); img2 = img2.contiguous().float()
    B, CH, H, W = img1.shape
    ssim_map = torch.empty_like(img1)
    if train:
        d1, d2, d3 = torch.empty_like(img1), torch.empty_like(img1), torch.empty_like(img1)
    else:
        d1 = d2 = d3 = None
    _L.check(L.ssb_fused_ssim_forward(C.c_int(B), C.c_int(CH), C.c_int(H), C.c_int(W), C.c_float(C1), C.c_float(C2),
                                      _L.ptr(img1), _L.ptr(img2), _L.ptr(ssim_map), _L.ptr(d1), _L.ptr(d2), _L.ptr(d3),
                                      _L.current_stream()), "ssb_fused_ssim_forward")
    if not train:
        e = torch.empty(0)
        return ssim_map, e, e, e
    return ssim_map, d1, d2, d3


def fusedssim_backward(C1, C2, img1, img2, dL_dmap, dm_dmu1, dm_dsigma1_sq, dm_dsigma12):
    """fused_ssim_cuda.fusedssim_backward (submodules/fused-ssim/ssim.cu:406-444)."""
    L = _L.lib()
    img1 = img1.contiguous().float(); img2 = img2.contiguous().float()
    B, CH, H, W = img1.shape
    if dm_dmu1.numel() == 0:
        raise RuntimeError("fusedssim_backward needs the derivative maps: call fusedssim(..., train=True)")
    out = torch.empty_like(img1)
    _L.check(L.ssb_fused_ssim_backward(C.c_int(B), C.c_int(CH), C.c_int(H), C.c_int(W), C.c_float(C1), C.c_float(C2),
                                       _L.ptr(img1), _L.ptr(img2), _L.ptr(dL_dmap.contiguous().float()), _L.ptr(dm_dmu1),
                                       _L.ptr(dm_dsigma1_sq), _L.ptr(dm_dsigma12), _L.ptr(out), _L.current_stream()),
             "ssb_fused_ssim_backward")
    return out


class FusedSSIMMap(torch.autograd.Function):
    @staticmethod
    def forward(ctx, C1, C2, img1, img2, padding="same", train=True):
        ssim_map, dm_dmu1, dm_dsigma1_sq, dm_dsigma12 = fusedssim(C1, C2, img1, img2, train)
        if padding == "valid":
            ssim_map = ssim_map[:, :, 5:-5, 5:-5]
        ctx.save_for_backward(img1.detach(), img2, dm_dmu1, dm_dsigma1_sq, dm_dsigma12)
        ctx.C1, ctx.C2, ctx.padding = C1, C2, padding
        return ssim_map

    @staticmethod
    def backward(ctx, opt_grad):
        img1, img2, dm_dmu1, dm_dsigma1_sq, dm_dsigma12 = ctx.saved_tensors
        dL_dmap = opt_grad
        if ctx.padding == "valid":
            dL_dmap = torch.zeros_like(img1)
            dL_dmap[:, :, 5:-5, 5:-5] = opt_grad
        grad = fusedssim_backward(ctx.C1, ctx.C2, img1, img2, dL_dmap, dm_dmu1, dm_dsigma1_sq, dm_dsigma12)
        return None, None, grad, None, None, None


def fused_ssim(img1, img2, padding="same", train=True):
    C1 = 0.01 ** 2
    C2 = 0.03 ** 2
    assert padding in allowed_padding
    map = FusedSSIMMap.apply(C1, C2, img1, img2, padding, train)
    return map.mean()
