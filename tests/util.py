"""Shared helpers of the test-suite: small seeded cases the CPU oracle finishes in seconds."""
import os
import sys
from dataclasses import replace

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def small_config(cfg, factor=4):
    """Same rig, images and focal lengths divided by `factor` (keeps ragged H36M widths ragged)."""
    sizes = tuple((max(32, w // factor + (2 if i in (0, 3) and w % 100 == 2 else 0)), max(32, h // factor)) for i, (w, h) in enumerate(cfg.image_sizes))
    return replace(cfg, image_sizes=sizes, focal_range=(cfg.focal_range[0] / factor, cfg.focal_range[1] / factor))


def raster_case(cfg, seed=0, n_views=2, big=True):
    """Numpy inputs of one frame for the rasteriser: anisotropic, rotated, partially transparent Gaussians with
    NON one-hot features (the op is generic), overlapping in screen space so that depth order matters."""
    from skelsplat_b200 import synthetic
    rng = np.random.default_rng(seed)
    seq = synthetic.make_sequence(cfg, 1, seed=seed)
    J = cfg.n_joints
    means = seq.frames[0].pose_3d_init.astype(np.float32)
    means[1] = means[0] + np.array([15, -10, 5], np.float32)          # near-coincident pair: overlapping footprints
    scales = np.exp(rng.uniform(3.5, 5.0, (J, 3)) if big else rng.uniform(2.5, 3.5, (J, 3))).astype(np.float32)
    q = rng.normal(size=(J, 4)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    opac = rng.uniform(0.3, 1.0, J).astype(np.float32)
    feats = (np.eye(J) + 0.1 * rng.uniform(size=(J, J))).astype(np.float32)
    cams = seq.cameras[:n_views]
    return dict(
        means3D=means, scales=scales, rotations=q, opacities=opac, features=feats,
        viewmatrix=np.stack([c.world_view_transform for c in cams]), projmatrix=np.stack([c.full_proj_transform for c in cams]),
        campos=np.stack([c.camera_center for c in cams]),
        dims=np.array([[c.image_width, c.image_height] for c in cams], np.int32),
        tanfov=np.array([[c.tanfovx, c.tanfovy] for c in cams], np.float32))


def relerr(a, b):
    """max|a-b| / max|b|: the tolerance form used for fp32 images and gradients (sums of mixed-sign terms)."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    if a.size == 0:
        return 0.0
    s = np.abs(b).max()
    d = np.abs(a - b).max()
    return float(d / s) if s > 0 else float(d)


def golden_path(name):
    return os.path.join(GOLDEN, name)


def have_golden(name):
    return os.path.exists(golden_path(name))


def synthetic_dL(shape, seed=0, scale=1e-3):
    """Deterministic dense 'upstream gradient' (a closed form, so fixtures need not store it)."""
    c, h, w = shape
    cc, yy, xx = np.meshgrid(np.arange(c), np.arange(h), np.arange(w), indexing="ij")
    return (scale * np.sin(0.37 * xx + 0.23 * yy + 1.3 * cc + seed) * np.cos(0.011 * xx * (seed + 1) - 0.017 * yy)).astype(np.float32)
