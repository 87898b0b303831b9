"""DLT multi-view triangulation: the producer of the initial guess (host side, numpy).

Same algorithm as the reference's triangulation.py:122-150 -- per joint stack the rows
``x*P[2]-P[0]`` and ``y*P[2]-P[1]`` of every view into A (2V x 4), take the right
singular vector of the smallest singular value and dehomogenise -- but batched over
joints (one ``numpy.linalg.svd`` call on a [J, 2V, 4] stack instead of J calls).
"""
import numpy as np


def build_dlt_rows(P_list, poses_2d):
    """A: [J, 2V, 4] for detections poses_2d [V, J, 2]."""
    P = np.asarray(P_list, np.float64)                     # [V,3,4]
    x = np.asarray(poses_2d, np.float64)                   # [V,J,2]
    rx = x[:, :, 0, None] * P[:, None, 2, :] - P[:, None, 0, :]   # [V,J,4]
    ry = x[:, :, 1, None] * P[:, None, 2, :] - P[:, None, 1, :]
    A = np.stack([rx, ry], axis=1)                         # [V,2,J,4]: rows ordered view-major, x then y
    return A.transpose(2, 0, 1, 3).reshape(x.shape[1], -1, 4)


def triangulate_poses(P_list, poses_2d):
    """[J,3] world points from V projection matrices K[R|t] and [V,J,2] detections."""
    A = build_dlt_rows(P_list, poses_2d)
    _, _, Vt = np.linalg.svd(A)
    X = Vt[:, -1, :]
    return X[:, :3] / X[:, 3:4]
