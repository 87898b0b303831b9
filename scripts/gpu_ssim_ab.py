"""Developer tool: time the fused-SSIM library surface (B=5, CH=1, 1500x1500) with the library SKELSPLAT_B200_LIB points at, and
print its deviation from a float64 conv2d SSIM on a small ragged shape.  Used to A/B compile-time variants of csrc/ssim.cu
(skelsplat_b200.build.build(defines=..., out=...))."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
import torch
from fused_ssim import fused_ssim


def ev(fn, reps=30, warm=5):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def ssim64(x, y):
    g = torch.tensor([np.exp(-(i - 5) ** 2 / (2 * 1.5 ** 2)) for i in range(11)], dtype=torch.float64, device=x.device); g = g / g.sum()
    w = (g[:, None] * g[None, :]).expand(x.shape[1], 1, 11, 11).contiguous()
    F = torch.nn.functional
    mu1, mu2 = F.conv2d(x, w, padding=5, groups=x.shape[1]), F.conv2d(y, w, padding=5, groups=x.shape[1])
    s1 = F.conv2d(x * x, w, padding=5, groups=x.shape[1]) - mu1 ** 2; s2 = F.conv2d(y * y, w, padding=5, groups=x.shape[1]) - mu2 ** 2
    s12 = F.conv2d(x * y, w, padding=5, groups=x.shape[1]) - mu1 * mu2
    return (((2 * mu1 * mu2 + 0.01 ** 2) * (2 * s12 + 0.03 ** 2)) / ((mu1 ** 2 + mu2 ** 2 + 0.01 ** 2) * (s1 + s2 + 0.03 ** 2))).mean()


torch.manual_seed(0)
a = torch.rand(5, 1, 1500, 1500, device="cuda").requires_grad_(True); b = torch.rand(5, 1, 1500, 1500, device="cuda")
def train():
    a.grad = None
    fused_ssim(a, b).backward()
out = {"lib": os.path.basename(os.environ.get("SKELSPLAT_B200_LIB", "default")), "inference_ms": round(ev(lambda: fused_ssim(a.detach(), b, train=False)), 4),
       "train_fwd_ms": round(ev(lambda: fused_ssim(a, b)), 4), "train_iter_ms": round(ev(train), 4)}
x = torch.rand(2, 3, 257, 301, device="cuda").requires_grad_(True); y = torch.rand(2, 3, 257, 301, device="cuda")
v = fused_ssim(x, y); v.backward()
x64 = x.detach().double().requires_grad_(True); r = ssim64(x64, y.double()); r.backward()
out["value_err"] = float((v.double() - r).abs()); out["grad_relerr"] = float((x.grad.double() - x64.grad).abs().max() / x64.grad.abs().max())
out["value"] = float(v); out["grad_checksum"] = float(x.grad.double().abs().sum())
print(json.dumps(out))
