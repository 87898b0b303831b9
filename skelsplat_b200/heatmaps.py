"""Pseudo-ground-truth heatmaps as ROI patches (per-frame setup feeding the hot path).

The reference builds, per frame, V dense ``[J,H,W]`` float32 heatmaps with one
cupy ``gaussian_filter`` call per (view, joint) on a megapixel image
(utils/general_utils.py:175-304).  Only a ``(2r_y+1) x (2r_x+1)`` window around each
detection is non-zero, so this module produces exactly those windows ("ROIs") and the
fused optimiser consumes them directly; ``rois_to_dense`` re-creates the dense tensor
for the drop-in (dense) API and for the parity tests.

Semantics kept literally from the reference:
  * sigma from the *initial* Gaussians' covariance pushed through ``T = R_w2c @ J`` with J
    row-wise -- which is NOT the rasteriser's EWA form (general_utils.py:224-246),
    +0.3 dilation, ``lambda = mid +- sqrt(max(0.1, mid^2 - det))``, ``sigma_y = sqrt(lambda1)`` on
    axis 0 and ``sigma_x = sqrt(lambda2)`` on axis 1 (general_utils.py:248-265);
  * peak at ``(clamp(int(v)), clamp(int(u)))`` with value 255 (general_utils.py:275-284);
  * separable Gaussian, ``truncate=4`` => half-width ``int(4*sigma+0.5)``, ``mode='reflect'``
    (cupyx/scipy ``gaussian_filter`` defaults, general_utils.py:289);
  * per-channel min-max normalisation with the +1e-8 (general_utils.py:300-304).
The filter arithmetic follows scipy.ndimage (float64 accumulation, float32 storage
between the two passes); cupy's float32 accumulation is not reproducible here --
"parity unpinned" for this setup step, as SURVEY.md section 8c records.  Two decisions
make the setup exactly reproducible between this specification and the GPU kernels
(csrc/setup.cu): sigma is evaluated in float64 with a fixed operation order
(``heatmap_sigmas``), so every integer window agrees; and a patch is stored FACTORED,
as its column and row profile, the heatmap value being their single fp32 product
(``HeatmapROIs``) -- the form the fused optimiser reads.
"""
from dataclasses import dataclass

import numpy as np


def _rotation_matrices(q):
    """build_rotation, utils/general_utils.py:87-108 (normalises q)."""
    q = q / np.sqrt((q * q).sum(1, keepdims=True))
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.zeros((q.shape[0], 3, 3), np.float32)
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - r * z); R[:, 0, 2] = 2 * (x * z + r * y)
    R[:, 1, 0] = 2 * (x * y + r * z); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - r * x)
    R[:, 2, 0] = 2 * (x * z - r * y); R[:, 2, 1] = 2 * (y * z + r * x); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def covariance_3d(scaling_raw, rotation, scaling_modifier=1.0):
    """GaussianModel.get_covariance -> unpack_covariance: Sigma = (R S)(R S)^T
    (scene/gaussian_model.py:33-37, utils/general_utils.py:110-119,144-165)."""
    s = np.exp(scaling_raw.astype(np.float32)) * np.float32(scaling_modifier)
    R = _rotation_matrices(rotation.astype(np.float32))
    L = R * s[:, None, :]
    return (L @ L.transpose(0, 2, 1)).astype(np.float32)


def heatmap_sigmas(xyz, cams, scaling_raw, rotation, scaling_modifier=1.0):
    """(sigma_y, sigma_x), each [V,J] float32 -- general_utils.py:199-265.

    The reference evaluates this with fp32 torch ops on its GPU; the window half-width ``int(4*sigma + 0.5)`` is a step
    function of sigma, so two fp32 evaluations that differ in the last bit disagree on ~3 % of the windows.  The
    specification is therefore written in float64 with a FIXED operation order (plain IEEE add/mul/div/sqrt on numpy
    float64 arrays: no matmul, no FMA) and rounded to fp32 at the end, which is what ``.item()`` of the reference's fp32
    tensor hands to the filter; csrc/setup.cu:roi_rect_kernel performs the same operations in the same order, so host and
    GPU agree on every window (tests/test_gpu_setup.py asserts 100 %)."""
    f64 = np.float64
    V, J = len(cams), xyz.shape[0]
    m = xyz.astype(np.float32).astype(f64)
    q = rotation.astype(np.float32).astype(f64)
    sr = scaling_raw.astype(np.float32).astype(f64)
    mod = f64(np.float32(scaling_modifier))
    qn = np.sqrt(((q[:, 0] * q[:, 0] + q[:, 1] * q[:, 1]) + q[:, 2] * q[:, 2]) + q[:, 3] * q[:, 3])
    r, x, y, z = q[:, 0] / qn, q[:, 1] / qn, q[:, 2] / qn, q[:, 3] / qn
    R = [[1.0 - 2.0 * (y * y + z * z), 2.0 * (x * y - r * z), 2.0 * (x * z + r * y)],
         [2.0 * (x * y + r * z), 1.0 - 2.0 * (x * x + z * z), 2.0 * (y * z - r * x)],
         [2.0 * (x * z - r * y), 2.0 * (y * z + r * x), 1.0 - 2.0 * (x * x + y * y)]]
    sk = [np.exp(sr[:, k]) * mod for k in range(3)]
    M = [[R[a][k] * sk[k] for k in range(3)] for a in range(3)]
    Sg = [[(M[a][0] * M[b][0] + M[a][1] * M[b][1]) + M[a][2] * M[b][2] for b in range(3)] for a in range(3)]
    s1 = np.zeros((V, J), np.float32); s2 = np.zeros((V, J), np.float32)
    for v, cam in enumerate(cams):
        vm = cam.world_view_transform.astype(np.float32).reshape(16).astype(f64)     # W2C(r, c) = vm[4 c + r]
        tfx, tfy = f64(np.float32(cam.tanfovx)), f64(np.float32(cam.tanfovy))
        t = [((vm[a] * m[:, 0] + vm[4 + a] * m[:, 1]) + vm[8 + a] * m[:, 2]) + vm[12 + a] for a in range(3)]
        fx, fy = f64(cam.image_width) / (2.0 * tfx), f64(cam.image_height) / (2.0 * tfy)
        limx, limy = 1.3 * tfx, 1.3 * tfy
        t[0] = np.minimum(limx, np.maximum(-limx, t[0] / t[2])) * t[2]
        t[1] = np.minimum(limy, np.maximum(-limy, t[1] / t[2])) * t[2]
        tz2 = t[2] * t[2]
        zero = np.zeros(J, f64)
        Jr = [[fx / t[2], zero, -((fx * t[0]) / tz2)], [zero, fy / t[2], -((fy * t[1]) / tz2)]]
        T = [[vm[a] * Jr[0][b] + vm[4 + a] * Jr[1][b] for b in range(3)] for a in range(3)]
        c00 = np.zeros(J, f64); c01 = np.zeros(J, f64); c11 = np.zeros(J, f64)
        for a in range(3):
            for b in range(3):
                c00 = c00 + (T[a][0] * Sg[b][a]) * T[b][0]
                c01 = c01 + (T[a][0] * Sg[b][a]) * T[b][1]
                c11 = c11 + (T[a][1] * Sg[b][a]) * T[b][1]
        cx, cz = c00 + 0.3, c11 + 0.3
        det = cx * cz - c01 * c01
        mid = 0.5 * (cx + cz)
        root = np.sqrt(np.maximum(0.1, mid * mid - det))
        s1[v] = np.sqrt(mid + root).astype(np.float32)
        s2[v] = np.sqrt(mid - root).astype(np.float32)
    return s1, s2


def _gauss_weights(sigma, radius):
    """scipy.ndimage._gaussian_kernel1d (order 0)."""
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    return phi / phi.sum()


def _filtered_delta_1d(n, pos, sigma, amplitude, dtype):
    """Response of scipy's ``correlate1d(mode='reflect')`` with a truncated Gaussian to a delta of
    ``amplitude`` at ``pos`` on an axis of length ``n``; returns (start, values) of its support."""
    radius = int(4.0 * float(sigma) + 0.5)
    w = _gauss_weights(float(sigma), radius)
    lo, hi = max(0, pos - radius), min(n - 1, pos + radius)
    idx = np.arange(lo, hi + 1)
    out = np.zeros(idx.shape[0], np.float64)
    # out[i] = sum_k w[k] * in[reflect(i + k - radius)]; 'reflect' = (d c b a | a b c d | d c b a)
    for k in range(2 * radius + 1):
        src = idx + k - radius
        src = np.where(src < 0, -src - 1, src)
        src = np.where(src >= n, 2 * n - 1 - src, src)
        out += np.where(src == pos, w[k] * amplitude, 0.0)
    return lo, out.astype(dtype)


@dataclass
class HeatmapROIs:
    """Packed ROI patches of one frame in FACTORED form.  rect[v,j] = (x0, y0, w, h).  Inside its window the heatmap of
    (v,j) is the product of a column profile and a row profile (a separable filter applied to a delta, then a scalar
    normalisation), so a patch is stored as ``h + w`` floats ``data[offset[v,j] : offset[v,j] + h + w]`` = col[h] | row[w]
    and the heatmap value at window pixel (a, b) is the ONE fp32 product ``col[a] * row[b]`` -- by definition: the dense
    tensors (``rois_to_dense``), the fused optimiser (csrc/optimizer.cu) and the reference arm all use exactly that value."""
    rect: np.ndarray      # [V,J,4] int32
    offset: np.ndarray    # [V,J]   int64
    data: np.ndarray      # [total] float32: per patch col[h] | row[w]
    sizes: list           # [(W,H)] per view

    def factors(self, v, j):
        x0, y0, w, h = self.rect[v, j]
        o = self.offset[v, j]
        return self.data[o:o + h], self.data[o + h:o + h + w]

    def patch(self, v, j):
        col, row = self.factors(v, j)
        return col[:, None] * row[None, :]            # float32 x float32 -> float32: one rounding per pixel


def heatmap_roi_rects(xyz_init, poses_2d, cams, scaling_raw, rotation, scaling_modifier=1.0):
    """Only the integer windows (x0, y0, w, h) [V,J,4] of generate_heatmap_rois -- the part that must agree EXACTLY between
    this specification and csrc/setup.cu:roi_rect_kernel (cheap enough to check on thousands of frames)."""
    V, J = len(cams), xyz_init.shape[0]
    poses_2d = np.asarray(poses_2d, np.float32)                    # the detections are an fp32 tensor when .long() truncates them
    s1, s2 = heatmap_sigmas(xyz_init, cams, scaling_raw, rotation, scaling_modifier)
    rect = np.zeros((V, J, 4), np.int32)
    for v, cam in enumerate(cams):
        W, H = cam.image_width, cam.image_height
        for j in range(J):
            xc = int(np.clip(int(poses_2d[v, j, 0]), 0, W - 1)); yc = int(np.clip(int(poses_2d[v, j, 1]), 0, H - 1))
            ry, rx = int(4.0 * float(s1[v, j]) + 0.5), int(4.0 * float(s2[v, j]) + 0.5)
            x0, x1, y0, y1 = max(0, xc - rx), min(W - 1, xc + rx), max(0, yc - ry), min(H - 1, yc + ry)
            rect[v, j] = (x0, y0, x1 - x0 + 1, y1 - y0 + 1)
    return rect


def generate_heatmap_rois(xyz_init, poses_2d, cams, scaling_raw, rotation, scaling_modifier=1.0):
    """ROI form of generate_heatmaps (utils/general_utils.py:175-297)."""
    f32 = np.float32
    V, J = len(cams), xyz_init.shape[0]
    poses_2d = np.asarray(poses_2d, np.float32)                    # the detections are an fp32 tensor when .long() truncates them
    s1, s2 = heatmap_sigmas(xyz_init, cams, scaling_raw, rotation, scaling_modifier)
    rect = np.zeros((V, J, 4), np.int32); offset = np.zeros((V, J), np.int64)
    chunks, total = [], 0
    for v, cam in enumerate(cams):
        W, H = cam.image_width, cam.image_height
        for j in range(J):
            xc = int(np.clip(int(poses_2d[v, j, 0]), 0, W - 1))    # .long() truncates toward zero
            yc = int(np.clip(int(poses_2d[v, j, 1]), 0, H - 1))
            y0, col = _filtered_delta_1d(H, yc, s1[v, j], 255.0, f32)      # pass 1 (axis 0), stored float32
            x0, row = _filtered_delta_1d(W, xc, s2[v, j], 1.0, np.float64)  # pass 2 weights
            # the filtered image is fl32(col[a] * row[b]); its maximum (rounding is monotonic) and the min-max normalisation
            # (x - 0) / (max - 0 + 1e-8) of general_utils.py:300-304 (the window never covers the whole image, so min = 0):
            mx = f32(np.float64(col.max()) * row.max())
            denom = f32(mx + f32(1e-8))
            # factored storage: the scalar goes into the row profile; value(a, b) := fl32(col[a] * rown[b]), which differs
            # from fl32(fl32(col[a] * row[b]) / denom) by at most ~1.5 ulp (the reference's own cupy filter differs by more)
            rown = (row / np.float64(denom)).astype(f32)
            rect[v, j] = (x0, y0, rown.shape[0], col.shape[0])
            offset[v, j] = total
            chunks.append(col.astype(f32)); chunks.append(rown)
            total += col.shape[0] + rown.shape[0]
    return HeatmapROIs(rect=rect, offset=offset, data=np.concatenate(chunks), sizes=[(c.image_width, c.image_height) for c in cams])


def rois_to_dense(rois: HeatmapROIs, v):
    """Dense ``[J,H,W]`` float32 heatmap of view v, as the reference stores it."""
    W, H = rois.sizes[v]
    J = rois.rect.shape[1]
    out = np.zeros((J, H, W), np.float32)
    for j in range(J):
        x0, y0, w, h = rois.rect[v, j]
        out[j, y0:y0 + h, x0:x0 + w] = rois.patch(v, j)
    return out
