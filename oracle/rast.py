"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of the C oracle (oracle/rast_oracle.c).

Restates, on the CPU, what ``CudaRasterizer::Rasterizer::forward/backward`` do
(RAST/cuda_rasterizer/rasterizer_impl.cu:198-450) and exposes every intermediate
stage (per-Gaussian state, unsorted/sorted keys, ranges) so that the CUDA product
can be checked stage by stage.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / reference legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    """Compile the C restatement (gcc, a second or two)."""
    subprocess.run(["make", "-s", "-C", _HERE, "oracle"], check=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "librast_oracle.so")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(os.path.join(_HERE, "rast_oracle.c")):
            build()
        _LIB = C.CDLL(path)
        _LIB.oracle_bin.restype = C.c_int
        _LIB.oracle_higher_msb.restype = C.c_uint32
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def forward(means3D, scales, rotations, opacities, features, viewmatrix, projmatrix, W, H,
            tanfovx, tanfovy, scale_modifier=1.0, cov3D_precomp=None, antialiasing=False):
    """Full forward.  features: [P, C].  viewmatrix/projmatrix: the 16 floats in the
    reference's memory order (torch row-major of the transposed matrix,
    scene/cameras.py:94-99).  Returns a dict with every stage's output."""
    L = lib()
    means3D = _f32(means3D); P = means3D.shape[0]
    scales = _f32(scales); rotations = _f32(rotations)
    opacities = _f32(opacities).reshape(-1); features = _f32(features)
    Cch = features.shape[1]
    view = _f32(viewmatrix).reshape(-1); proj = _f32(projmatrix).reshape(-1)
    cov_pre = _f32(cov3D_precomp)
    gx, gy = (W + 15) // 16, (H + 15) // 16
    radii = np.zeros(P, np.int32); means2D = np.zeros((P, 2), np.float32)
    depths = np.zeros(P, np.float32); cov3D = np.zeros((P, 6), np.float32)
    conic_opacity = np.zeros((P, 4), np.float32); tiles_touched = np.zeros(P, np.uint32)
    rects = np.zeros((P, 4), np.uint32)
    L.oracle_preprocess(C.c_int(P), _p(means3D), _p(scales), C.c_float(scale_modifier), _p(rotations),
                        _p(opacities), _p(cov_pre), _p(view), _p(proj), C.c_int(W), C.c_int(H),
                        C.c_float(tanfovx), C.c_float(tanfovy), C.c_int(int(antialiasing)),
                        _p(radii), _p(means2D), _p(depths), _p(cov3D), _p(conic_opacity),
                        _p(tiles_touched), _p(rects))
    if cov_pre is not None:
        cov3D = cov_pre.reshape(P, 6).copy()
    R = int(tiles_touched.sum())
    n = max(R, 1)
    offsets = np.zeros(P, np.uint32)
    ku = np.zeros(n, np.uint64); vu = np.zeros(n, np.uint32)
    ks = np.zeros(n, np.uint64); vs = np.zeros(n, np.uint32)
    ranges = np.zeros((gx * gy, 2), np.uint32)
    R2 = L.oracle_bin(C.c_int(P), C.c_int(W), C.c_int(H), _p(radii), _p(means2D), _p(depths),
                      _p(tiles_touched), _p(offsets), _p(ku), _p(vu), _p(ks), _p(vs), _p(ranges))
    assert R2 == R
    final_T = np.zeros((H, W), np.float32); n_contrib = np.zeros((H, W), np.uint32)
    color = np.zeros((Cch, H, W), np.float32); invdepth = np.zeros((1, H, W), np.float32)
    L.oracle_render_forward(C.c_int(Cch), C.c_int(W), C.c_int(H), _p(ranges), _p(vs), _p(means2D),
                            _p(features), _p(conic_opacity), _p(depths), _p(final_T), _p(n_contrib),
                            _p(color), _p(invdepth))
    return dict(R=R, color=color, invdepth=invdepth, radii=radii, means2D=means2D, depths=depths,
                cov3D=cov3D, conic_opacity=conic_opacity, tiles_touched=tiles_touched, rects=rects,
                point_offsets=offsets, keys_unsorted=ku[:R], vals_unsorted=vu[:R],
                keys_sorted=ks[:R], point_list=vs[:R], ranges=ranges, final_T=final_T,
                n_contrib=n_contrib)


def backward(fwd, means3D, scales, rotations, features, viewmatrix, projmatrix, W, H, tanfovx, tanfovy,
             dL_dcolor, dL_dinvdepth=None, scale_modifier=1.0):
    """Full backward given the dict returned by forward().  Returns the gradient
    tensors of RasterizeGaussiansBackwardCUDA (RAST/rasterize_points.cu:126-223) except dL_dsh."""
    L = lib()
    means3D = _f32(means3D); P = means3D.shape[0]
    scales = _f32(scales); rotations = _f32(rotations); features = _f32(features)
    Cch = features.shape[1]
    view = _f32(viewmatrix).reshape(-1); proj = _f32(projmatrix).reshape(-1)
    dL_dcolor = _f32(dL_dcolor); dL_dinv = _f32(dL_dinvdepth)
    dmean2D = np.zeros((P, 3), np.float32); dconic = np.zeros((P, 2, 2), np.float32)
    dopacity = np.zeros((P, 1), np.float32); dcolors = np.zeros((P, Cch), np.float32)
    dinvdepths = np.zeros((P, 1), np.float32)
    L.oracle_render_backward(C.c_int(P), C.c_int(Cch), C.c_int(W), C.c_int(H), _p(fwd["ranges"]),
                             _p(fwd["point_list"] if fwd["R"] else np.zeros(1, np.uint32)),
                             _p(fwd["means2D"]), _p(fwd["conic_opacity"]), _p(features), _p(fwd["depths"]),
                             _p(fwd["final_T"]), _p(fwd["n_contrib"]), _p(dL_dcolor), _p(dL_dinv),
                             _p(dmean2D), _p(dconic), _p(dopacity), _p(dcolors),
                             _p(dinvdepths) if dL_dinv is not None else None)
    dmean3D = np.zeros((P, 3), np.float32); dcov3D = np.zeros((P, 6), np.float32)
    dscale = np.zeros((P, 3), np.float32); drot = np.zeros((P, 4), np.float32)
    L.oracle_preprocess_backward(C.c_int(P), _p(means3D), _p(fwd["radii"]), _p(fwd["cov3D"]), _p(scales),
                                 _p(rotations), C.c_float(scale_modifier), _p(view), _p(proj),
                                 C.c_int(W), C.c_int(H), C.c_float(tanfovx), C.c_float(tanfovy),
                                 _p(dmean2D), _p(dconic), _p(dinvdepths) if dL_dinv is not None else None,
                                 _p(dmean3D), _p(dcov3D), _p(dscale), _p(drot))
    return dict(dL_dmeans2D=dmean2D, dL_dcolors=dcolors, dL_dopacity=dopacity, dL_dmeans3D=dmean3D,
                dL_dcov3D=dcov3D, dL_dscales=dscale, dL_drotations=drot, dL_dconic=dconic)


def mark_visible(means3D, viewmatrix):
    L = lib()
    means3D = _f32(means3D); P = means3D.shape[0]
    out = np.zeros(P, np.uint8)
    L.oracle_mark_visible(C.c_int(P), _p(means3D), _p(_f32(viewmatrix).reshape(-1)), _p(out))
    return out.astype(bool)


def higher_msb(n):
    return int(lib().oracle_higher_msb(C.c_uint32(n)))
