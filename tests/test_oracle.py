"""The C oracle against (a) itself (structural properties of keys / sort / ranges), (b) an independent float64
autograd restatement of the forward math (validates the analytic backward chain), (c) the golden fixtures
generated from the UNMODIFIED reference kernels (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import rast, torch_renderer
from skelsplat_b200 import configs
from tests.util import small_config, raster_case, relerr, golden_path, have_golden, synthetic_dL


def run_oracle(case, vi, features=None):
    W, H = int(case["dims"][vi, 0]), int(case["dims"][vi, 1])
    f = rast.forward(case["means3D"], case["scales"], case["rotations"], case["opacities"],
                     case["features"] if features is None else features, case["viewmatrix"][vi], case["projmatrix"][vi],
                     W, H, float(case["tanfov"][vi, 0]), float(case["tanfov"][vi, 1]))
    return f, W, H


@pytest.mark.parametrize("name", ["h36m", "panoptic", "occlusion-person"])
def test_binning_structure(name):
    cfg = small_config(configs.get_config(name))
    case = raster_case(cfg, seed=3)
    f, W, H = run_oracle(case, 0)
    gx = (W + 15) // 16
    R = f["R"]
    assert R == int(f["tiles_touched"].sum()) == int(f["point_offsets"][-1]) and R > 0
    # emission order: Gaussian-major, row-major tiles inside the rect
    k = 0
    for i in range(cfg.n_joints):
        x0, y0, x1, y1 = f["rects"][i]
        for y in range(y0, y1):
            for x in range(x0, x1):
                assert f["vals_unsorted"][k] == i
                assert f["keys_unsorted"][k] >> np.uint64(32) == y * gx + x
                assert np.uint32(f["keys_unsorted"][k] & np.uint64(0xFFFFFFFF)) == f["depths"][i:i + 1].view(np.uint32)[0]
                k += 1
    assert k == R
    # sorted keys ascending, same multiset, stable
    ks = f["keys_sorted"]
    assert np.all(ks[1:] >= ks[:-1])
    assert sorted(zip(f["keys_unsorted"].tolist(), f["vals_unsorted"].tolist())) == sorted(zip(ks.tolist(), f["point_list"].tolist()))
    # ranges partition [0,R) by tile, untouched tiles are (0,0)
    tiles = (ks >> np.uint64(32)).astype(np.int64)
    covered = 0
    for t in range(f["ranges"].shape[0]):
        s, e = f["ranges"][t]
        if e > s:
            assert np.all(tiles[s:e] == t) and (s == 0 or tiles[s - 1] != t) and (e == R or tiles[e] != t)
            covered += e - s
        else:
            assert (s, e) == (0, 0) and not np.any(tiles == t)
    assert covered == R


def test_higher_msb():
    for n, want in ((3969, 12), (8160, 13), (3600, 12), (1, 1), (4096, 13), (255, 8)):
        assert rast.higher_msb(n) == want


def test_render_matches_float64_restatement_and_autograd_matches_analytic_backward():
    cfg = small_config(configs.H36M, factor=8)
    case = raster_case(cfg, seed=5)
    f, W, H = run_oracle(case, 0)
    t64 = lambda a, g=False: torch.tensor(np.asarray(a, np.float64), requires_grad=g)
    m, s, q, o = t64(case["means3D"], True), t64(case["scales"], True), t64(case["rotations"], True), t64(case["opacities"], True)
    color, invd = torch_renderer.render(m, s, q, o, t64(case["features"]), t64(case["viewmatrix"][0]), t64(case["projmatrix"][0]),
                                        W, H, float(case["tanfov"][0, 0]), float(case["tanfov"][0, 1]))
    # borderline pixels may flip a threshold between fp32 and fp64: compare robustly
    diff = np.abs(color.detach().numpy() - f["color"])
    assert np.quantile(diff, 0.999) < 2e-5 and (diff > 1e-3).mean() < 1e-4
    rng = np.random.default_rng(0)
    dL = rng.normal(size=f["color"].shape).astype(np.float32) * (f["color"] > 0)
    dLi = rng.normal(size=f["invdepth"].shape).astype(np.float32) * 10.0
    (color * t64(dL)).sum().backward(retain_graph=True)
    (invd * t64(dLi)).sum().backward()
    g = rast.backward(f, case["means3D"], case["scales"], case["rotations"], case["features"], case["viewmatrix"][0],
                      case["projmatrix"][0], W, H, float(case["tanfov"][0, 0]), float(case["tanfov"][0, 1]), dL, dLi)
    assert relerr(g["dL_dmeans3D"], m.grad.numpy()) < 2e-3
    assert relerr(g["dL_dscales"], s.grad.numpy()) < 2e-3
    assert relerr(g["dL_drotations"], q.grad.numpy()) < 2e-3
    assert relerr(g["dL_dopacity"].reshape(-1), o.grad.numpy()) < 2e-3


def test_culled_and_degenerate_gaussians():
    cfg = small_config(configs.H36M)
    case = raster_case(cfg, seed=7)
    case["means3D"][2] = case["campos"][0] - 1000 * np.array(case["viewmatrix"][0][:3, 2])   # behind the camera
    case["means3D"][3] += 1e6                                                                  # far off-screen
    f, W, H = run_oracle(case, 0)
    assert f["radii"][2] == 0 and f["tiles_touched"][2] == 0
    assert f["tiles_touched"][3] == 0 and f["radii"][3] == 0
    vis = rast.mark_visible(case["means3D"], case["viewmatrix"][0])
    assert not vis[2] and vis[0]
    assert 2 not in f["point_list"] and 3 not in f["point_list"]
    dL = np.ones_like(f["color"])
    g = rast.backward(f, case["means3D"], case["scales"], case["rotations"], case["features"], case["viewmatrix"][0],
                      case["projmatrix"][0], W, H, float(case["tanfov"][0, 0]), float(case["tanfov"][0, 1]), dL)
    assert not g["dL_dmeans3D"][2].any() and not g["dL_dscales"][3].any()


@pytest.mark.parametrize("variant", ["h36m", "panoptic", "op"])
def test_oracle_against_reference_golden(variant):
    """Pins the oracle: every stage of the C restatement vs the outputs of the reference's own kernels."""
    if not have_golden(f"raster_{variant}.npz"):
        pytest.skip("golden fixture not generated yet (tests/golden/make_golden.py on a GPU box)")
    G = np.load(golden_path(f"raster_{variant}.npz"))
    case = {k: G[k] for k in ("means3D", "scales", "rotations", "opacities", "features", "viewmatrix", "projmatrix", "campos", "dims", "tanfov")}
    for vi in range(case["viewmatrix"].shape[0]):
        f, W, H = run_oracle(case, vi)
        p = f"v{vi}_"
        assert f["R"] == int(G[p + "R"])
        for k in ("radii", "tiles_touched", "point_offsets", "keys_unsorted", "vals_unsorted", "keys_sorted", "point_list", "ranges", "n_contrib"):
            assert np.array_equal(f[k], G[p + k]), k                               # integer / index work: bit-exact
        vis = G[p + "radii"] > 0
        for k in ("depths", "means2D", "conic_opacity", "cov3D"):
            assert np.array_equal(f[k][vis].view(np.uint32), G[p + k][vis].view(np.uint32)), k   # fp32 state feeding the keys: bit-exact
        assert relerr(f["color"], G[p + "color"]) < 1e-5                           # glibc expf vs CUDA expf
        # and element-wise, not only max-normalised: every pixel above 1e-3 within 1e-5 of its own value, the same non-zero set
        ref_c, my_c = G[p + "color"].astype(np.float64), f["color"].astype(np.float64)
        big = np.abs(ref_c) > 1e-3
        assert (np.abs(my_c - ref_c)[big] / np.abs(ref_c)[big]).max() <= 1e-5
        assert np.abs(my_c - ref_c)[~big].max() <= 3e-8
        assert np.array_equal(my_c != 0, ref_c != 0)
        assert relerr(f["invdepth"], G[p + "invdepth"]) < 1e-5
        assert relerr(f["final_T"], G[p + "final_T"]) < 1e-5
        g = rast.backward(f, case["means3D"], case["scales"], case["rotations"], case["features"], case["viewmatrix"][vi],
                          case["projmatrix"][vi], W, H, float(case["tanfov"][vi, 0]), float(case["tanfov"][vi, 1]), synthetic_dL(G[p + "color"].shape, vi), synthetic_dL(G[p + "invdepth"].shape, 10 + vi))
        for k in ("dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dscales", "dL_drotations"):
            spread = relerr(G[p + k + "_run2"], G[p + k])                           # the reference's own atomics noise
            assert relerr(g[k].reshape(-1), G[p + k].reshape(-1)) < max(2e-5, 4 * spread), k
