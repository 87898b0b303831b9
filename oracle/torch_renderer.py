"""TEST INFRASTRUCTURE ONLY -- a second, independent oracle: the rasteriser's MATH in float64 torch
with autograd, used to validate the analytic backward chain of oracle/rast_oracle.c (and hence of the
CUDA kernels) without trusting any hand-written gradient formula.

It restates the forward semantics only (RAST/cuda_rasterizer/forward.cu:74-401): EWA projection with the
1.3*tanfov clamp and +0.3 dilation, tile-rectangle gating, per-pixel front-to-back compositing in global
depth order with the alpha<1/255, power>0 and T<1e-4 rules.  It is not bit-exact (float64, vectorised)
and is meant for small images.
"""
import torch


def render(means3D, scales, rotations, opacities, features, viewmatrix, projmatrix, W, H, tanfovx, tanfovy, scale_modifier=1.0):
    """All inputs float64 tensors (requires_grad as wanted).  viewmatrix/projmatrix: [4,4] in the reference's
    memory order (= transposed true matrices).  Returns (color [C,H,W], invdepth [1,H,W])."""
    dt = torch.float64
    P = means3D.shape[0]
    view = viewmatrix.to(dt).T          # true W2C
    proj = projmatrix.to(dt).T          # true full projection
    hom = torch.cat([means3D, torch.ones(P, 1, dtype=dt)], 1)
    t = (view @ hom.T).T[:, :3]
    ph = (proj @ hom.T).T
    p_w = 1.0 / (ph[:, 3] + 1e-7)
    ndc = ph[:, :2] * p_w[:, None]
    focal_x, focal_y = W / (2.0 * tanfovx), H / (2.0 * tanfovy)
    # cov3D = (S R)^T (S R) in the reference's storage == R S S R^T with R built from the UN-normalised quaternion
    r, x, y, z = rotations[:, 0], rotations[:, 1], rotations[:, 2], rotations[:, 3]
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                     2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                     2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], 1).reshape(P, 3, 3)
    L = R @ torch.diag_embed(scale_modifier * scales)
    Sigma = L @ L.transpose(1, 2)
    limx, limy = 1.3 * tanfovx, 1.3 * tanfovy
    tz = t[:, 2]
    tx = torch.clamp(t[:, 0] / tz, -limx, limx) * tz
    ty = torch.clamp(t[:, 1] / tz, -limy, limy) * tz
    zero = torch.zeros_like(tz)
    J = torch.stack([focal_x / tz, zero, -(focal_x * tx) / (tz * tz),
                     zero, focal_y / tz, -(focal_y * ty) / (tz * tz)], 1).reshape(P, 2, 3)
    Wm = view[:3, :3]
    M = J @ Wm                           # 2x3
    cov = M @ Sigma @ M.transpose(1, 2)
    cxx, cxy, cyy = cov[:, 0, 0] + 0.3, cov[:, 0, 1], cov[:, 1, 1] + 0.3
    det = cxx * cyy - cxy * cxy
    conx, cony, conz = cyy / det, -cxy / det, cxx / det
    mid = 0.5 * (cxx + cyy)
    lam = mid + torch.sqrt(torch.clamp(mid * mid - det, min=0.1))
    radius = torch.ceil(3.0 * torch.sqrt(lam)).detach()
    px = ((ndc[:, 0] + 1.0) * W - 1.0) * 0.5
    py = ((ndc[:, 1] + 1.0) * H - 1.0) * 0.5
    gx, gy = (W + 15) // 16, (H + 15) // 16
    ys, xs = torch.meshgrid(torch.arange(H, dtype=dt), torch.arange(W, dtype=dt), indexing="ij")
    tile_x, tile_y = (xs // 16), (ys // 16)
    order = sorted(range(P), key=lambda i: (float(tz[i]), i))
    T = torch.ones(H, W, dtype=dt)
    done = torch.zeros(H, W, dtype=torch.bool)
    C = features.shape[1]
    color = torch.zeros(C, H, W, dtype=dt)
    invd = torch.zeros(H, W, dtype=dt)
    for i in order:
        if float(tz[i]) <= 0.2 or float(det[i]) == 0.0:
            continue
        rad = float(radius[i]); cx_, cy_ = float(px[i]), float(py[i])
        x0 = min(gx, max(0, int((cx_ - rad) / 16))); y0 = min(gy, max(0, int((cy_ - rad) / 16)))
        x1 = min(gx, max(0, int((cx_ + rad + 15) / 16))); y1 = min(gy, max(0, int((cy_ + rad + 15) / 16)))
        if (x1 - x0) * (y1 - y0) == 0:
            continue
        in_rect = (tile_x >= x0) & (tile_x < x1) & (tile_y >= y0) & (tile_y < y1)
        dx, dy = px[i] - xs, py[i] - ys
        power = -0.5 * (conx[i] * dx * dx + conz[i] * dy * dy) - cony[i] * dx * dy
        alpha = torch.clamp(opacities[i] * torch.exp(power), max=0.99)
        ok = in_rect & (power <= 0) & (alpha >= 1.0 / 255.0) & ~done
        test_T = T * (1 - alpha)
        newly_done = ok & (test_T < 1e-4)
        acc = ok & ~newly_done
        w = torch.where(acc, alpha * T, torch.zeros_like(T))
        color = color + features[i].to(dt)[:, None, None] * w[None]
        invd = invd + w / tz[i]
        T = torch.where(acc, test_T, T)
        done = done | newly_done
    return color, invd[None]
