"""TEST INFRASTRUCTURE (not imported by the product).  Host mirror (numpy, the readable specification) of the sort-free binning of the fused optimiser (optimizer.cu, phase B).

The reference bins by sorting (tile << 32 | depth bits) keys with a stable radix sort (rasterizer_impl.cu:70-111, 303-320).
Every Gaussian's tiles form a rectangle and there are at most 20 Gaussians, so the position of a (Gaussian j, tile t) pair in
that sorted list has a closed form:

    pos(j, t) = sum over j' of #{tiles of rect(j') that precede t in row-major order}
              + #{j' : t in rect(j') and (depth bits, id)(j') < (depth bits, id)(j)}

(the second term is the stable sort's tie order: equal tiles are ordered by depth, equal depths by emission order = id).
The CUDA kernel evaluates this with one thread per pair; tests/test_host_logic.py checks the formula against the C oracle's
sort on random scenes."""
import numpy as np


def sorted_positions(rects, depth_bits):
    """rects [P,4] (x0, y0, x1, y1) exclusive maxima in tile units (all-zero / empty for culled Gaussians); depth_bits [P] uint32.
    Returns (gaussian_of_pair [R], tile_xy_of_pair [R,2], pos_of_pair [R]) in EMISSION order (Gaussian-major, row-major in its
    rectangle), i.e. pos_of_pair[i] is where emitted pair i lands in the reference's sorted list."""
    rects = np.asarray(rects, np.int64)
    depth_bits = np.asarray(depth_bits, np.uint64)
    P = rects.shape[0]
    order = np.lexsort((np.arange(P), depth_bits))        # (depth bits, id) order
    rank = np.empty(P, np.int64)
    rank[order] = np.arange(P)
    gs, xy, pos = [], [], []
    for j in range(P):
        x0, y0, x1, y1 = rects[j]
        for y in range(y0, y1):
            for x in range(x0, x1):
                p = 0
                for o in range(P):
                    ox0, oy0, ox1, oy1 = rects[o]
                    ow = ox1 - ox0
                    p += min(max(y - oy0, 0), oy1 - oy0) * ow                      # whole rows of rect(o) above y
                    if oy0 <= y < oy1:
                        p += min(max(x - ox0, 0), ow)                              # same row, left of x
                        if ox0 <= x < ox1 and rank[o] < rank[j]:
                            p += 1                                                  # same tile, nearer Gaussian
                gs.append(j); xy.append((x, y)); pos.append(p)
    return np.asarray(gs, np.int64), np.asarray(xy, np.int64).reshape(-1, 2), np.asarray(pos, np.int64)


def row_band(py, conx, cony, conz):
    """Pixel rows [rlo, rhi] outside which a splat cannot reach alpha >= 1/255 (optimizer.cu, phase A): for a fixed dy the largest
    power over dx is -0.5 dy^2 det/conx, which must be >= -5.55 (pair_alpha's early-out threshold; opacity <= 1), so
    |dy| <= sqrt(11.1 conx/det), widened by 1 % + 1 px against fp32 rounding; a near-singular conic gets no band."""
    f = np.float32
    py, conx, cony, conz = f(py), f(conx), f(cony), f(conz)
    cdet = f(f(conx * conz) - f(cony * cony))
    if conx > 0 and cdet > f(1e-4) * conx * conz:
        ey = f(1.01) * np.sqrt(f(11.1) * conx / cdet, dtype=np.float32) + f(1.0)
    else:
        ey = f(1e9)
    rlo = int(min(max(np.ceil(py - ey), 0.0), 65535.0))
    rhi = int(min(max(np.floor(py + ey), -1.0), 65535.0))
    return rlo, rhi
