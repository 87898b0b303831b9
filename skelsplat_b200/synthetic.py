"""Synthetic H36M / Panoptic / Occlusion-Person-shaped inputs.

There is no dataset in this environment, so the benchmark and the parity tests run
on seeded synthetic sequences: random skeletons from a kinematic tree, a ring of
calibrated cameras looking at the subject, noisy 2D detections, and the DLT initial
guess the reference computes with triangulation.py.  All units are millimetres.

What a "frame" carries is exactly what the reference's DataLoader yields per scene
(scene/dataset_readers.py:84-238): ``(pose_3d_init, pose_3d_gt, poses_2d[V,J,2], cameras)``.
"""
from dataclasses import dataclass
from typing import List

import numpy as np

from .cameras import ViewCamera, make_camera
from .configs import SceneConfig
from .triangulation import triangulate_poses

# Kinematic trees.  Joint order follows each dataset's convention so that the
# limb pairs of limb_3d_consistency_loss (utils/loss_utils.py:226-250) are real limbs.
_H36M_PARENTS = [-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 9, 8, 11, 12, 8, 14, 15]
_H36M_LEN = [0, 132, 443, 454, 132, 443, 454, 233, 257, 121, 115, 151, 279, 252, 151, 279, 252]
_H36M_DIR = [(0, 0, 0), (-1, 0, 0), (0, 0, -1), (0, 0, -1), (1, 0, 0), (0, 0, -1), (0, 0, -1), (0, 0, 1),
             (0, 0, 1), (0, 0, 1), (0, 0, 1), (1, 0, 0), (0, 0, -1), (0, 0, -1), (-1, 0, 0), (0, 0, -1), (0, 0, -1)]
_H36M_MIRROR = {1: 4, 2: 5, 3: 6, 14: 11, 15: 12, 16: 13}

_PAN_PARENTS = [2, 0, -1, 0, 3, 4, 2, 6, 7, 0, 9, 10, 2, 12, 13, 1, 1, 15, 16]
_PAN_LEN = [480, 180, 0, 170, 280, 250, 110, 420, 430, 170, 280, 250, 110, 420, 430, 40, 40, 80, 80]
_PAN_DIR = [(0, 0, 1), (0, 0.5, 1), (0, 0, 0), (1, 0, 0), (0, 0, -1), (0, 0, -1), (1, 0, -0.3), (0, 0, -1), (0, 0, -1),
            (-1, 0, 0), (0, 0, -1), (0, 0, -1), (-1, 0, -0.3), (0, 0, -1), (0, 0, -1), (-0.6, 0.2, 0.6), (0.6, 0.2, 0.6),
            (-1, -0.6, 0), (1, -0.6, 0)]
_PAN_MIRROR = {9: 3, 10: 4, 11: 5, 12: 6, 13: 7, 14: 8, 15: 16, 17: 18}

_OP_PARENTS = [-1, 0, 1, 2, 0, 4, 5, 0, 7, 7, 9, 10, 7, 12, 13]
_OP_LEN = [0, 130, 440, 450, 130, 440, 450, 500, 200, 160, 280, 250, 160, 280, 250]
_OP_DIR = [(0, 0, 0), (-1, 0, 0), (0, 0, -1), (0, 0, -1), (1, 0, 0), (0, 0, -1), (0, 0, -1), (0, 0, 1), (0, 0, 1),
           (1, 0, 0), (0, 0, -1), (0, 0, -1), (-1, 0, 0), (0, 0, -1), (0, 0, -1)]
_OP_MIRROR = {1: 4, 2: 5, 3: 6, 12: 9, 13: 10, 14: 11}

_TREES = {
    17: (_H36M_PARENTS, _H36M_LEN, _H36M_DIR, _H36M_MIRROR),
    19: (_PAN_PARENTS, _PAN_LEN, _PAN_DIR, _PAN_MIRROR),
    15: (_OP_PARENTS, _OP_LEN, _OP_DIR, _OP_MIRROR),
}


@dataclass
class Frame:
    pose_3d_init: np.ndarray   # [J,3] float64  initial guess (DLT of the noisy detections)
    pose_3d_gt: np.ndarray     # [J,3] float64
    poses_2d: np.ndarray       # [V,J,2] float64 detections (pixels)
    name: str


@dataclass
class Sequence:
    cfg: SceneConfig
    cameras: List[ViewCamera]
    frames: List[Frame]


def _rot(axis, ang):
    axis = np.asarray(axis, float)
    axis = axis / (np.linalg.norm(axis) + 1e-12)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * (K @ K)


def random_skeleton(rng, n_joints):
    """Random pose on a fixed kinematic tree; left/right limb lengths are equal in GT
    (the reference's limb-consistency prior assumes it)."""
    parents, lens, dirs, mirror = _TREES[n_joints]
    scale = rng.uniform(0.9, 1.1, size=n_joints)
    for a, b in mirror.items():
        scale[a] = scale[b]
    root = np.array([rng.uniform(-500, 500), rng.uniform(-500, 500), rng.uniform(800, 1100)])
    yaw = _rot((0, 0, 1), rng.uniform(-np.pi, np.pi))
    X = np.zeros((n_joints, 3))
    order = sorted(range(n_joints), key=lambda j: _depth(parents, j))
    for j in order:
        p = parents[j]
        if p < 0:
            X[j] = root
            continue
        d = np.asarray(dirs[j], float)
        d = d / np.linalg.norm(d)
        wob = _rot(rng.normal(size=3), rng.uniform(0, 0.6))
        X[j] = X[p] + yaw @ (wob @ d) * lens[j] * scale[j]
    return X


def _depth(parents, j):
    d = 0
    while parents[j] >= 0:
        j = parents[j]
        d += 1
    return d


def make_cameras(rng, cfg: SceneConfig):
    """V cameras on a ring, looking at the origin +-200 mm, OpenCV convention (z forward, y down)."""
    cams = []
    V = cfg.nviews
    phase = rng.uniform(0, 2 * np.pi)
    for v in range(V):
        W, H = cfg.image_sizes[v % len(cfg.image_sizes)]
        ang = phase + 2 * np.pi * v / V + rng.uniform(-0.2, 0.2)
        rad = rng.uniform(*cfg.cam_ring_radius_mm)
        c = np.array([rad * np.cos(ang), rad * np.sin(ang), rng.uniform(1200, 2500)])
        target = np.array([rng.uniform(-200, 200), rng.uniform(-200, 200), 900 + rng.uniform(-200, 200)])
        z = target - c
        z /= np.linalg.norm(z)
        x = np.cross(z, np.array([0.0, 0.0, 1.0]))
        x /= np.linalg.norm(x)
        y = np.cross(z, x)
        R = np.stack([x, y, z])            # rows = camera axes in world coords -> X_cam = R (X - c)
        t = -R @ c
        f = rng.uniform(*cfg.focal_range)
        fy = f * rng.uniform(0.998, 1.002)
        K = np.array([[f, 0, W / 2 + rng.uniform(-15, 15)], [0, fy, H / 2 + rng.uniform(-15, 15)], [0, 0, 1.0]])
        cams.append(make_camera(v, K, R, t, W, H))
    return cams


def project(cam: ViewCamera, X):
    Xc = X @ cam.R_w2c.T + cam.t
    uv = Xc @ cam.K.T
    return uv[:, :2] / uv[:, 2:3]


def make_sequence(cfg: SceneConfig, n_frames: int, seed: int = 0, shard: int = None) -> Sequence:
    """``shard`` (optional): the camera rig is that of ``seed`` but the frames come from an independent stream keyed by
    (seed, shard) -- the shards of ONE sequence (one rig) that the ranks of a multi-GPU run optimise."""
    rng = np.random.default_rng(seed)
    cams = make_cameras(rng, cfg)
    if shard is not None:
        rng = np.random.default_rng([seed, 7919 + shard])
    frames = []
    P_list = [c.P3x4() for c in cams]
    for f in range(n_frames):
        gt = random_skeleton(rng, cfg.n_joints)
        det = np.stack([project(c, gt) for c in cams])
        det = det + rng.normal(0.0, cfg.det_noise_px, size=det.shape)
        if cfg.occluded:
            # h36m-occ: 2-3 joints in 1-2 views replaced by gross outliers
            for v in rng.choice(cfg.nviews, size=rng.integers(1, 3), replace=False):
                for j in rng.choice(cfg.n_joints, size=rng.integers(2, 4), replace=False):
                    det[v, j] += rng.normal(0.0, 40.0, size=2)
        init = triangulate_poses(P_list, det)
        frames.append(Frame(pose_3d_init=init, pose_3d_gt=gt, poses_2d=det, name=f"S1_Synth_{f}"))
    return Sequence(cfg=cfg, cameras=cams, frames=frames)
