"""Camera matrices in the reference's conventions (host side, numpy).

Mirrors, without the per-frame disk I/O, what the reference builds per frame:
  * getWorld2View2           utils/graphics_utils.py:38-49
  * getProjectionMatrix2     utils/graphics_utils.py:74-95
  * focal2fov                utils/graphics_utils.py:100-101
  * Camera.__init__          scene/cameras.py:88-100  (world_view_transform = W2C^T,
                             full_proj_transform = W2C^T @ P^T, camera_center)
  * getNerfppNorm            scene/dataset_readers.py:482-503 (cameras_extent -> spatial_lr_scale)

The kernels read the 16 floats of ``world_view_transform`` / ``full_proj_transform``
in torch row-major order, i.e. element (r, c) of the true matrix sits at
``m[4*c + r]`` (RAST/cuda_rasterizer/auxiliary.h:70-89).
"""
import math
from dataclasses import dataclass

import numpy as np

ZNEAR, ZFAR = 0.01, 100.0  # scene/cameras.py:88-89


def focal2fov(focal, pixels):
    return 2 * math.atan(pixels / (2 * focal))


def world2view(R_w2c, t):
    """4x4 float32 W2C.  The reference passes R = R_w2c^T and transposes it back
    (getWorld2View2 with translate=0, scale=1, including its inverse/inverse round trip)."""
    Rt = np.zeros((4, 4))
    Rt[:3, :3] = R_w2c
    Rt[:3, 3] = t
    Rt[3, 3] = 1.0
    C2W = np.linalg.inv(Rt)
    Rt = np.linalg.inv(C2W)
    return np.float32(Rt)


def projection_matrix2(K, W, H, znear=ZNEAR, zfar=ZFAR):
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    top = znear * cy / fy
    bottom = -znear * (H - cy) / fy
    right = znear * (W - cx) / fx
    left = -znear * cx / fx
    P = np.zeros((4, 4), np.float32)  # torch.zeros(4, 4) in the reference: float32 storage
    P[0, 0] = 2.0 * znear / (right - left)
    P[1, 1] = 2.0 * znear / (top - bottom)
    P[0, 2] = -(right + left) / (right - left)
    P[1, 2] = (top + bottom) / (top - bottom)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


@dataclass
class ViewCamera:
    """One calibrated view, carrying exactly what render_* reads from a reference ``Camera``."""
    uid: int
    image_width: int
    image_height: int
    K: np.ndarray            # 3x3 float64
    R_w2c: np.ndarray        # 3x3 float64, X_cam = R X + t
    t: np.ndarray            # 3   float64
    FoVx: float
    FoVy: float
    world_view_transform: np.ndarray   # 4x4 float32 = W2C^T
    projection_matrix: np.ndarray      # 4x4 float32 = P^T
    full_proj_transform: np.ndarray    # 4x4 float32
    camera_center: np.ndarray          # 3 float32

    @property
    def tanfovx(self):
        return math.tan(self.FoVx * 0.5)

    @property
    def tanfovy(self):
        return math.tan(self.FoVy * 0.5)

    def P3x4(self):
        """K [R|t], as triangulation.py:59-67 builds it."""
        return self.K @ np.hstack([self.R_w2c, self.t.reshape(3, 1)])


def make_camera(uid, K, R_w2c, t, W, H):
    K = np.asarray(K, np.float64)
    R_w2c = np.asarray(R_w2c, np.float64)
    t = np.asarray(t, np.float64).reshape(3)
    wvt = world2view(R_w2c, t).T.copy()
    proj = projection_matrix2(K, W, H).T.copy()
    full = (wvt @ proj).astype(np.float32)
    center = np.linalg.inv(wvt)[3, :3].astype(np.float32)
    return ViewCamera(uid=uid, image_width=int(W), image_height=int(H), K=K, R_w2c=R_w2c, t=t,
                      FoVx=focal2fov(K[0, 0], W), FoVy=focal2fov(K[1, 1], H),
                      world_view_transform=wvt, projection_matrix=proj, full_proj_transform=full,
                      camera_center=center)


def cameras_extent(cams):
    """1.1 * max_v || c_v - mean(c) ||  (getNerfppNorm); becomes ``spatial_lr_scale``."""
    centers = []
    for cam in cams:
        W2C = world2view(cam.R_w2c, cam.t)
        centers.append(np.linalg.inv(W2C)[:3, 3:4])
    centers = np.hstack(centers)
    avg = np.mean(centers, axis=1, keepdims=True)
    dist = np.linalg.norm(centers - avg, axis=0, keepdims=True)
    return float(np.max(dist) * 1.1)
