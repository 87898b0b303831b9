"""Result sink: in-memory MPJPE exactly as eval.py defines it, and writers for the
on-disk artefacts the reference's own eval.py / initial-guess tooling expect.

Reference: eval.py:90-170 walks ``experiments/<run>/point_cloud/iteration_500/<scene>.ply`` with open3d, computes the
absolute error ``mean_j ||pred_j - gt_j||`` (eval.py:122-123) and the root-relative error after subtracting joint 0
(eval.py:137-138), per action and overall; scene/gaussian_model.py:264-281 writes one PLY per frame via plyfile.
Here the poses never leave memory: errors are computed on the ``[F,J,3]`` array, and ``write_ply`` / ``save_poses_npz``
emit the same information without plyfile/open3d (binary little-endian PLY with the reference's vertex attributes).
"""
import os
from collections import defaultdict

import numpy as np


def mpjpe_absolute(pred, gt):
    """Per-frame absolute MPJPE in mm: mean over joints of the Euclidean distance (eval.py:122-123)."""
    return np.linalg.norm(np.asarray(pred, np.float64) - np.asarray(gt, np.float64), axis=-1).mean(axis=-1)


def mpjpe_root_relative(pred, gt, root=0):
    """Per-frame root-relative MPJPE: both poses translated so that joint ``root`` is the origin (eval.py:131-138)."""
    pred = np.asarray(pred, np.float64); gt = np.asarray(gt, np.float64)
    return mpjpe_absolute(pred - pred[..., root:root + 1, :], gt - gt[..., root:root + 1, :])


def evaluate(pred, gt, scene_names=None):
    """Overall and per-action table like eval.py:140-170.  Scene names follow '<subject>_<action>_<step>'."""
    abs_e, rel_e = mpjpe_absolute(pred, gt), mpjpe_root_relative(pred, gt)
    out = {"absolute_mpjpe_mm": float(abs_e.mean()), "relative_mpjpe_mm": float(rel_e.mean()), "n_frames": int(abs_e.shape[0])}
    if scene_names is not None:
        per = defaultdict(list)
        for n, a, r in zip(scene_names, abs_e, rel_e):
            parts = n.split("_")
            per[parts[1] if len(parts) > 1 else n].append((a, r))
        out["per_action"] = {k: {"absolute_mpjpe_mm": float(np.mean([x[0] for x in v])),
                                 "relative_mpjpe_mm": float(np.mean([x[1] for x in v])), "n": len(v)} for k, v in sorted(per.items())}
    return out


def write_ply(path, xyz, features_dc=None, opacity=None, scaling=None, rotation=None):
    """One frame's Gaussians as a binary PLY with the attribute list of GaussianModel.construct_list_of_attributes
    (scene/gaussian_model.py:250-281): x y z nx ny nz f_dc_* opacity scale_* rot_*.  With only ``xyz`` given it is
    the plain point cloud triangulation.py:195-200 writes (x y z)."""
    xyz = np.asarray(xyz, np.float32)
    J = xyz.shape[0]
    cols, names = [xyz], ["x", "y", "z"]
    if features_dc is not None:
        f = np.asarray(features_dc, np.float32).reshape(J, -1)
        cols = [xyz, np.zeros((J, 3), np.float32), f]
        names += ["nx", "ny", "nz"] + [f"f_dc_{i}" for i in range(f.shape[1])]
        cols.append(np.asarray(opacity, np.float32).reshape(J, 1)); names.append("opacity")
        sc = np.asarray(scaling, np.float32).reshape(J, -1); cols.append(sc); names += [f"scale_{i}" for i in range(sc.shape[1])]
        ro = np.asarray(rotation, np.float32).reshape(J, -1); cols.append(ro); names += [f"rot_{i}" for i in range(ro.shape[1])]
    data = np.ascontiguousarray(np.concatenate(cols, axis=1), dtype="<f4")
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    header = "ply\nformat binary_little_endian 1.0\nelement vertex %d\n" % J + "".join(f"property float {n}\n" for n in names) + "end_header\n"
    with open(path, "wb") as fh:
        fh.write(header.encode("ascii"))
        fh.write(data.tobytes())


def read_ply_xyz(path):
    """Minimal reader for the files written above (binary little-endian, float properties)."""
    with open(path, "rb") as fh:
        raw = fh.read()
    end = raw.index(b"end_header\n") + len(b"end_header\n")
    lines = raw[:end].decode("ascii").splitlines()
    n = int([l for l in lines if l.startswith("element vertex")][0].split()[-1])
    props = [l.split()[-1] for l in lines if l.startswith("property float")]
    arr = np.frombuffer(raw[end:end + 4 * n * len(props)], dtype="<f4").reshape(n, len(props))
    return arr[:, [props.index("x"), props.index("y"), props.index("z")]].copy()


def save_poses_npz(path, xyz, scene_names=None, gt=None):
    """All final poses of a run in one file (replaces one PLY round trip per frame)."""
    extra = {}
    if scene_names is not None:
        extra["scene_names"] = np.asarray(scene_names)
    if gt is not None:
        extra["gt"] = np.asarray(gt, np.float32)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    np.savez_compressed(path, xyz=np.asarray(xyz, np.float32), **extra)
