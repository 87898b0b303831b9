"""TEST INFRASTRUCTURE (not imported by the product).  Host mirror (numpy, the readable specification) of what one view-iteration of the fused optimiser computes per tile
(optimizer.cu, phase C + the first lines of phase D): forward replay, the one-hot scalar recurrence of the backward, the raw
moment sums per Gaussian, the loss-mask size N and the mapping of the sums to dL/dmean2D, dL/dconic, dL/dopacity.

Derivation.  Features are one-hot (Gaussian j renders only into channel j), the loss is l2_gaussian,
L = sum_mask (render - gt)^2 / N with mask = (gt > 0) | (render > 0).  At a pixel, for the Gaussians of its tile in depth order,
render_j = alpha_j T_j.  The reference's backward (backward.cu:560-600) keeps per CHANNEL  accum_rec[ch] <- last_alpha last_color[ch]
+ (1 - last_alpha) accum_rec[ch]  and  dL/dalpha_j = sum_ch (c_j[ch] - accum_rec[ch]) dL/dpix[ch] T_j; with one-hot colours only
the channels of the Gaussians BEHIND j (and j's own) are non-zero in that sum, and it collapses to ONE scalar recurrence
    S <- alpha_last g_last + (1 - alpha_last) S,      dL/dalpha_j = (g_j - S) T_j,      g_j = dL/dpix[channel j] = 2 err_j / N.
Everything that is constant per Gaussian is pulled out of the pixel loop: with w = G (err - S) T the kernel accumulates
    sum w dx, sum w dy, sum w dx^2, sum w dx dy, sum w dy^2, sum w        (dx = mean_x - pixel_x, ...)
and applies opacity, conic, the factor 2, -1/2, the NDC scale and 1/N once (records_to_grads)."""
import numpy as np

ALPHA_MAX, ALPHA_MIN, T_EPS = np.float32(0.99), np.float32(1.0 / 255.0), np.float32(1e-4)


def view_iteration(means2D, conic_opacity, ranges, point_list, gt, W, H):
    """means2D [P,2], conic_opacity [P,4], ranges [tiles,2], point_list [R] from the binning; gt [P,H,W] (channel j = Gaussian j).
    Returns dict(dL_dmean2D [P,2], dL_dconic [P,3] (xx, xy, yy), dL_dopacity [P], loss, N, render [P,H,W])."""
    f = np.float32
    P = means2D.shape[0]
    gx = (W + 15) // 16
    r = np.zeros((P, 8), np.float64)
    render = np.zeros((P, H, W), np.float32)
    for tile in np.nonzero(ranges[:, 1] > ranges[:, 0])[0]:
        ty, tx = divmod(int(tile), gx)
        lst = [int(g) for g in point_list[ranges[tile, 0]:ranges[tile, 1]]]
        for py in range(ty * 16, min(ty * 16 + 16, H)):
            for px in range(tx * 16, min(tx * 16 + 16, W)):
                T = f(1.0)
                contrib = []
                for g in lst:                                               # forward.cu:330-386
                    dx = f(means2D[g, 0] - f(px)); dy = f(means2D[g, 1] - f(py))
                    cx, cy, cz, op = (f(v) for v in conic_opacity[g])
                    power = f(f(-0.5) * f(f(cx * dx * dx) + f(cz * dy * dy)) - f(cy * dx * dy))
                    if power > 0:
                        continue
                    G = f(np.exp(power))
                    alpha = min(ALPHA_MAX, f(op * G))
                    if alpha < ALPHA_MIN:
                        continue
                    test_T = f(T * f(1 - alpha))
                    if test_T < T_EPS:
                        break
                    contrib.append((g, alpha, G, T, dx, dy))
                    render[g, py, px] = alpha * T
                    T = test_T
                S = la = lg = 0.0
                for g, alpha, G, Tb, dx, dy in reversed(contrib):          # backward.cu:536-636, one-hot form
                    gv = float(gt[g, py, px])
                    err = float(alpha) * float(Tb) - gv
                    S = la * lg + (1.0 - la) * S
                    lg, la = err, float(alpha)
                    w = float(G) * (err - S) * float(Tb)
                    r[g] += (w * dx, w * dy, w * dx * dx, w * dx * dy, w * dy * dy, w, err * err - max(gv, 0.0) ** 2, 0.0 if gv > 0 else 1.0)
    n_gt = int((gt > 0).sum())
    N = n_gt + r[:, 7].sum()
    loss = (r[:, 6].sum() + float((np.where(gt > 0, gt, 0).astype(np.float64) ** 2).sum())) / N
    op = conic_opacity[:, 3].astype(np.float64); cx, cy, cz = (conic_opacity[:, i].astype(np.float64) for i in range(3))
    t, u = 2 * op * r[:, 0], 2 * op * r[:, 1]
    return dict(dL_dmean2D=np.stack([-(cx * t + cy * u) * 0.5 * W, -(cz * u + cy * t) * 0.5 * H], 1) / N,
                dL_dconic=np.stack([-op * r[:, 2], -op * r[:, 3], -op * r[:, 4]], 1) / N, dL_dopacity=2 * r[:, 5] / N,
                loss=loss, N=N, render=render)


def adam_step_fp32(param, grad, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-15):
    """One torch.optim.Adam step exactly as the fused kernel performs it (optimizer.cu phase E + the host step table):
    python-float (fp64) host scalars step_size = lr / (1 - beta1^t) and sqrt(1 - beta2^t), rounded to fp32 where torch hands
    them to an fp32 tensor op; then lerp / mul+addcmul / sqrt / div / add / addcdiv in fp32.  Arrays are float32, updated
    copies are returned.  tests/test_host_logic.py checks bit-equality with torch.optim.Adam (foreach path)."""
    f = np.float32
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    neg_step = f(-(lr / bc1))
    bc2_sqrt = f(np.sqrt(bc2))
    m = (m + f(1.0 - beta1) * (grad - m)).astype(f)                                  # lerp_(grad, 1 - beta1)  (fma in the kernel)
    v = (f(beta2) * v).astype(f)
    v = (v + (f(1.0 - beta2) * grad).astype(f) * grad).astype(f)                      # addcmul_(grad, grad, value=1 - beta2)
    denom = ((np.sqrt(v).astype(f) / bc2_sqrt).astype(f) + f(eps)).astype(f)
    param = (param + neg_step * (m / denom).astype(f)).astype(f)                      # addcdiv_(m, denom, value=-step_size)
    return param, m, v
