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


class _FusedSSIMMean(torch.autograd.Function):
    """map.mean() fused into the kernels: the SSIM map and dL/dmap (a constant) never touch HBM.  Same value as
    FusedSSIMMap.apply(...).mean() up to the summation order (per-warp fp32 partials, summed in fp64 in a fixed order)."""

    @staticmethod
    def forward(ctx, C1, C2, img1, img2, padding, train):
        L = _L.lib()
        a = img1.contiguous().float(); b = img2.contiguous().float()
        B, CH, H, W = a.shape
        crop = 5 if padding == "valid" else 0
        out = torch.empty((), dtype=torch.float32, device=a.device)
        ws = torch.empty(int(L.ssb_fused_ssim_mean_workspace_bytes(C.c_int(B), C.c_int(CH), C.c_int(H), C.c_int(W))) // 4,
                         dtype=torch.float32, device=a.device)
        d1 = d2 = d3 = None
        if train:
            d1, d2, d3 = torch.empty_like(a), torch.empty_like(a), torch.empty_like(a)
        _L.check(L.ssb_fused_ssim_mean_forward(C.c_int(B), C.c_int(CH), C.c_int(H), C.c_int(W), C.c_float(C1), C.c_float(C2), _L.ptr(a), _L.ptr(b),
                                               C.c_int(crop), _L.ptr(out), _L.ptr(d1), _L.ptr(d2), _L.ptr(d3), _L.ptr(ws), _L.current_stream()),
                 "ssb_fused_ssim_mean_forward")
        ctx.crop, ctx.train = crop, train
        if train:
            ctx.save_for_backward(a.detach(), b, d1, d2, d3)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        if not ctx.train:
            raise RuntimeError("fused_ssim(..., train=False) does not keep the derivative maps: call it with train=True to differentiate")
        L = _L.lib()
        a, b, d1, d2, d3 = ctx.saved_tensors
        B, CH, H, W = a.shape
        g = grad_out.contiguous().float()
        grad = torch.empty_like(a)
        _L.check(L.ssb_fused_ssim_mean_backward(C.c_int(B), C.c_int(CH), C.c_int(H), C.c_int(W), _L.ptr(a), _L.ptr(b), _L.ptr(g), C.c_int(ctx.crop),
                                                _L.ptr(d1), _L.ptr(d2), _L.ptr(d3), _L.ptr(grad), _L.current_stream()), "ssb_fused_ssim_mean_backward")
        return None, None, grad, None, None, None


def fused_ssim(img1, img2, padding="same", train=True):
    C1 = 0.01 ** 2
    C2 = 0.03 ** 2
    assert padding in allowed_padding
    if img1.dim() != 4 or img1.shape[-1] <= 10 or img1.shape[-2] <= 10:      # degenerate sizes: the map path handles them like the reference
        return FusedSSIMMap.apply(C1, C2, img1, img2, padding, train).mean()
    return _FusedSSIMMean.apply(C1, C2, img1, img2, padding, train)
