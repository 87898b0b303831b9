"""Parity of the CUDA dense rasteriser (through the C ABI) with the C oracle, the golden fixtures of the
reference kernels and -- when oracle/_ref is present on the box -- the reference kernels themselves.
Bar: tile keys / sort order / ranges / radii / per-Gaussian state BIT-EXACT; image and gradients <= 1e-5
relative (fp32; relative to the tensor's max magnitude, gradients being sums of mixed-sign terms)."""
import numpy as np
import pytest
import torch

from oracle import rast
from skelsplat_b200 import configs, synthetic
from skelsplat_b200 import rasterizer as R
from tests.util import small_config, raster_case, relerr, golden_path, have_golden, synthetic_dL

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 1e-5
INT_KEYS = ("tiles_touched", "point_offsets", "keys_unsorted", "vals_unsorted", "keys_sorted", "point_list", "ranges")
F32_KEYS = ("depths", "means2D", "conic_opacity", "cov3D")


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def mine_forward(case, vi, **kw):
    W, H = int(case["dims"][vi, 0]), int(case["dims"][vi, 1])
    P = case["means3D"].shape[0]
    out = R.rasterize_batched(t(case["means3D"])[None], t(case["scales"])[None], t(case["rotations"])[None], t(case["opacities"]).reshape(1, P),
                              t(case["features"]), t(case["viewmatrix"][vi])[None], t(case["projmatrix"][vi])[None], W, H,
                              float(case["tanfov"][vi, 0]), float(case["tanfov"][vi, 1]), **kw)
    torch.cuda.synchronize()
    return out, W, H


def mine_backward(case, vi, st, W, H, dL, dLinv=None):
    P = case["means3D"].shape[0]
    g = R.rasterize_batched_backward(st, t(case["means3D"])[None], t(case["scales"])[None], t(case["rotations"])[None], t(case["opacities"]).reshape(1, P),
                                     t(case["features"]), t(case["viewmatrix"][vi])[None], t(case["projmatrix"][vi])[None], W, H,
                                     float(case["tanfov"][vi, 0]), float(case["tanfov"][vi, 1]), t(dL)[None], None if dLinv is None else t(dLinv)[None])
    torch.cuda.synchronize()
    return {k: v[0].cpu().numpy() for k, v in g.items()}


def oracle_forward(case, vi):
    W, H = int(case["dims"][vi, 0]), int(case["dims"][vi, 1])
    return rast.forward(case["means3D"], case["scales"], case["rotations"], case["opacities"], case["features"], case["viewmatrix"][vi],
                        case["projmatrix"][vi], W, H, float(case["tanfov"][vi, 0]), float(case["tanfov"][vi, 1]))


def assert_stages_equal(ms, ref, radii, ref_radii):
    assert ms["R"] == ref["R"]
    assert np.array_equal(radii, ref_radii)
    for k in INT_KEYS:
        assert np.array_equal(ms[k], ref[k]), k
    vis = ref_radii > 0
    for k in F32_KEYS:
        assert np.array_equal(ms[k][vis].view(np.uint32), ref[k][vis].view(np.uint32)), k


@pytest.mark.parametrize("name", ["h36m", "panoptic", "occlusion-person"])
@pytest.mark.parametrize("big", [True, False])
def test_forward_and_backward_vs_oracle(name, big):
    cfg = small_config(configs.get_config(name))
    case = raster_case(cfg, seed=21, big=big)
    for vi in range(2):
        (color, radii, invd, st), W, H = mine_forward(case, vi)
        of = oracle_forward(case, vi)
        assert_stages_equal(st.parse(0), of, radii[0].cpu().numpy(), of["radii"])
        assert relerr(color[0].cpu().numpy(), of["color"]) < TOL
        assert relerr(invd[0].cpu().numpy(), of["invdepth"]) < TOL
        rng = np.random.default_rng(vi)
        dL = (rng.normal(size=of["color"].shape) * 1e-3).astype(np.float32)
        dLinv = (rng.normal(size=of["invdepth"].shape) * 1e-3).astype(np.float32)
        g = mine_backward(case, vi, st, W, H, dL, dLinv)
        og = rast.backward(of, case["means3D"], case["scales"], case["rotations"], case["features"], case["viewmatrix"][vi], case["projmatrix"][vi],
                           W, H, float(case["tanfov"][vi, 0]), float(case["tanfov"][vi, 1]), dL, dLinv)
        for k, ok in (("means3D", "dL_dmeans3D"), ("means2D", "dL_dmeans2D"), ("scales", "dL_dscales"), ("rotations", "dL_drotations"),
                      ("opacity", "dL_dopacity"), ("features", "dL_dcolors"), ("cov3D", "dL_dcov3D"), ("conic", "dL_dconic")):
            assert relerr(g[k].reshape(-1), og[ok].reshape(-1)) < 2 * TOL, k


@pytest.mark.parametrize("variant", ["h36m", "panoptic", "op"])
def test_against_reference_golden(variant):
    if not have_golden(f"raster_{variant}.npz"):
        pytest.skip("golden fixture missing")
    G = np.load(golden_path(f"raster_{variant}.npz"))
    case = {k: G[k] for k in ("means3D", "scales", "rotations", "opacities", "features", "viewmatrix", "projmatrix", "campos", "dims", "tanfov")}
    for vi in range(case["viewmatrix"].shape[0]):
        (color, radii, invd, st), W, H = mine_forward(case, vi)
        p = f"v{vi}_"
        ref = {k: G[p + k] for k in INT_KEYS + F32_KEYS}
        ref["R"] = int(G[p + "R"])
        assert_stages_equal(st.parse(0), ref, radii[0].cpu().numpy(), G[p + "radii"])
        assert relerr(color[0].cpu().numpy(), G[p + "color"]) < TOL
        assert relerr(invd[0].cpu().numpy(), G[p + "invdepth"]) < TOL
        g = mine_backward(case, vi, st, W, H, synthetic_dL(G[p + "color"].shape, vi), synthetic_dL(G[p + "invdepth"].shape, 10 + vi))
        for k, rk in (("means3D", "dL_dmeans3D"), ("means2D", "dL_dmeans2D"), ("scales", "dL_dscales"), ("rotations", "dL_drotations"),
                      ("opacity", "dL_dopacity"), ("features", "dL_dcolors"), ("cov3D", "dL_dcov3D")):
            spread = relerr(G[p + rk + "_run2"], G[p + rk])          # the reference's own atomics noise
            assert relerr(g[k].reshape(-1), G[p + rk].reshape(-1)) < max(TOL, 4 * spread), k


@pytest.mark.parametrize("name,variant", [("h36m", "h36m"), ("panoptic", "panoptic"), ("occlusion-person", "op")])
def test_full_size_against_reference_kernels(name, variant):
    """BASELINE.json full sizes, against the UNMODIFIED reference kernels running beside us."""
    from oracle import ref_rasterizer as refr
    if not refr.available(variant):
        pytest.skip("oracle/_ref not present on this box")
    cfg = configs.get_config(name)
    case = raster_case(cfg, seed=31, n_views=cfg.nviews, big=False)
    e = torch.Tensor([]); bg = torch.zeros(32, device=DEV)
    J = cfg.n_joints
    for vi in range(cfg.nviews):
        (color, radii, invd, st), W, H = mine_forward(case, vi)
        Rn, rcolor, rradii, geom, binning, img, rinvd = refr.rasterize_forward(
            variant, bg, t(case["means3D"]), e, t(case["opacities"]).reshape(-1, 1), t(case["scales"]), t(case["rotations"]), 1.0, e,
            t(case["viewmatrix"][vi]), t(case["projmatrix"][vi]), float(case["tanfov"][vi, 0]), float(case["tanfov"][vi, 1]), H, W,
            t(case["features"]).reshape(J, 1, J), 0, t(case["campos"][vi]))
        rs = refr.RefState(geom, binning, img, Rn, J, W, H, variant).parse()
        assert_stages_equal(st.parse(0), rs, radii[0].cpu().numpy(), rradii.cpu().numpy())
        mine_img, ref_img = color[0].cpu().numpy(), rcolor.cpu().numpy()
        assert relerr(mine_img, ref_img) < TOL
        # ELEMENT-WISE (not max-normalised): same expf, same operation order as forward.cu:352-396 => every pixel within 1e-5 of
        # its own value where it is not tiny, the SAME set of non-zero pixels, exact zeros everywhere else
        assert_elementwise(mine_img, ref_img, "color")
        assert_elementwise(invd[0].cpu().numpy(), rinvd.cpu().numpy(), "invdepth")
        assert np.array_equal(mine_img != 0, ref_img != 0)
        # size-independent property: every element is written (no stale memory), untouched tiles are exactly zero
        assert torch.isfinite(color).all()
        # full-size BACKWARD against the reference kernels themselves (two reference runs give its own atomics spread)
        dL = synthetic_dL(ref_img.shape, seed=vi); dLinv = synthetic_dL((1, H, W), seed=10 + vi)
        g = mine_backward(case, vi, st, W, H, dL, dLinv)
        runs = []
        for _ in range(2):
            rg = refr.rasterize_backward(variant, bg, t(case["means3D"]), rradii, e, t(case["opacities"]).reshape(-1, 1), t(case["scales"]),
                                         t(case["rotations"]), 1.0, e, t(case["viewmatrix"][vi]), t(case["projmatrix"][vi]),
                                         float(case["tanfov"][vi, 0]), float(case["tanfov"][vi, 1]), t(dL), t(dLinv),
                                         t(case["features"]).reshape(J, 1, J), 0, t(case["campos"][vi]), geom, Rn, binning, img)
            runs.append([x.cpu().numpy() for x in rg])
        names = ("means2D", "features", "opacity", "means3D", "cov3D", None, "scales", "rotations")     # order of the reference's return tuple
        for k, a, b in zip(names, runs[0], runs[1]):
            if k is None:
                continue                                       # dL_dsh: garbage in the reference
            spread = relerr(b, a)
            assert relerr(g[k].reshape(-1), a.reshape(-1)) < max(TOL, 4 * spread), (k, relerr(g[k].reshape(-1), a.reshape(-1)), spread)


def assert_elementwise(a, b, what, rtol=1e-5, floor=1e-3, atol=2e-8):
    """|a - b| <= rtol |b| wherever |b| > floor, and <= atol + rtol*floor below it."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    big = np.abs(b) > floor
    if big.any():
        r = (np.abs(a - b)[big] / np.abs(b)[big]).max()
        assert r <= rtol, (what, "relative", r, "bit-equal fraction", float((a == b).mean()))
    if (~big).any():
        d = np.abs(a - b)[~big].max()
        assert d <= atol + rtol * floor, (what, "absolute below floor", d)


def test_batched_ragged_views_equal_single_view_calls():
    """frames x views batching with per-view (W,H) and packed outputs == one call per view, bit for bit."""
    cfg = small_config(configs.H36M)
    seq = synthetic.make_sequence(cfg, 3, seed=8)
    J, V, F = cfg.n_joints, cfg.nviews, 3
    rng = np.random.default_rng(0)
    means = np.stack([f.pose_3d_init for f in seq.frames]).astype(np.float32)
    scales = np.exp(rng.uniform(3.5, 4.8, (F, J, 3))).astype(np.float32)
    rots = rng.normal(size=(F, J, 4)).astype(np.float32); rots /= np.linalg.norm(rots, axis=-1, keepdims=True)
    opac = rng.uniform(0.4, 1, (F, J)).astype(np.float32)
    feats = np.eye(J, dtype=np.float32)
    cams = seq.cameras
    vm = np.stack([c.world_view_transform for c in cams]); pm = np.stack([c.full_proj_transform for c in cams])
    dims = np.array([[c.image_width, c.image_height] for c in cams], np.int32)
    tanfov = np.array([[c.tanfovx, c.tanfovy] for c in cams], np.float32)
    Wm, Hm = int(dims[:, 0].max()), int(dims[:, 1].max())
    sizes = [J * int(dims[b % V, 0]) * int(dims[b % V, 1]) for b in range(F * V)]
    coff = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64)
    isz = [int(dims[b % V, 0]) * int(dims[b % V, 1]) for b in range(F * V)]
    ioff = np.concatenate([[0], np.cumsum(isz)[:-1]]).astype(np.int64)
    out_color = torch.full((int(np.sum(sizes)),), float("nan"), device=DEV)
    out_inv = torch.full((int(np.sum(isz)),), float("nan"), device=DEV)
    _, radii, _, st = R.rasterize_batched(t(means), t(scales), t(rots), t(opac), t(feats), t(vm), t(pm), Wm, Hm, 0.0, 0.0,
                                          dims=t(dims), tanfov=t(tanfov), color_offsets=t(coff), invdepth_offsets=t(ioff),
                                          out_color=out_color, out_invdepth=out_inv)
    dL_all = torch.from_numpy((rng.normal(size=int(np.sum(sizes))) * 1e-3).astype(np.float32)).to(DEV)
    gb = R.rasterize_batched_backward(st, t(means), t(scales), t(rots), t(opac), t(feats), t(vm), t(pm), Wm, Hm, 0.0, 0.0, dL_all,
                                      dims=t(dims), tanfov=t(tanfov), color_offsets=t(coff))
    torch.cuda.synchronize()
    assert not torch.isnan(out_color).any() and not torch.isnan(out_inv).any()       # every element written
    for b in range(F * V):
        f, v = b // V, b % V
        W, H = int(dims[v, 0]), int(dims[v, 1])
        c1, r1, i1, st1 = R.rasterize_batched(t(means[f])[None], t(scales[f])[None], t(rots[f])[None], t(opac[f])[None], t(feats),
                                              t(vm[v])[None], t(pm[v])[None], W, H, float(tanfov[v, 0]), float(tanfov[v, 1]))
        assert torch.equal(c1.reshape(-1), out_color[coff[b]:coff[b] + sizes[b]])
        assert torch.equal(i1.reshape(-1), out_inv[ioff[b]:ioff[b] + isz[b]])
        assert torch.equal(r1[0], radii[b])
        g1 = R.rasterize_batched_backward(st1, t(means[f])[None], t(scales[f])[None], t(rots[f])[None], t(opac[f])[None], t(feats),
                                          t(vm[v])[None], t(pm[v])[None], W, H, float(tanfov[v, 0]), float(tanfov[v, 1]),
                                          dL_all[coff[b]:coff[b] + sizes[b]].reshape(1, J, H, W).contiguous())
        for k in ("means3D", "scales", "rotations", "opacity"):
            assert torch.equal(g1[k][0], gb[k][b]), k


def test_backward_is_deterministic():
    cfg = small_config(configs.PANOPTIC)
    case = raster_case(cfg, seed=4)
    (color, radii, invd, st), W, H = mine_forward(case, 0)
    dL = np.random.default_rng(1).normal(size=color[0].shape).astype(np.float32)
    a = mine_backward(case, 0, st, W, H, dL)
    b = mine_backward(case, 0, st, W, H, dL)
    for k in a:
        assert np.array_equal(a[k], b[k]), k


def test_edge_cases():
    cfg = small_config(configs.H36M)
    case = raster_case(cfg, seed=9)
    # (1) everything culled: behind the camera
    c2 = dict(case); c2["means3D"] = (case["campos"][0][None] - 1000 * case["viewmatrix"][0][:3, 2][None]).repeat(cfg.n_joints, 0).astype(np.float32)
    (color, radii, invd, st), W, H = mine_forward(c2, 0)
    assert st.num_rendered() == 0 and not radii.any() and not color.any() and not invd.any()
    g = mine_backward(c2, 0, st, W, H, np.ones((cfg.n_joints, H, W), np.float32))
    assert all(not v.any() for v in g.values())
    # (2) r_capacity overflow is flagged, never silent
    (color, radii, invd, st), W, H = mine_forward(case, 0, r_capacity=32)
    assert st.header()[2] == 1 and st.header()[7] > 32
    with pytest.raises(Exception):
        st.check()
    # (3) precomputed 3D covariance gives the same image as scales+rotations (the cov3D is what the state stores)
    (c_ref, _, _, st_ref), W, H = mine_forward(case, 0)
    cov = st_ref.parse()["cov3D"]
    P = cfg.n_joints
    out = R.rasterize_batched(t(case["means3D"])[None], None, None, t(case["opacities"]).reshape(1, P), t(case["features"]),
                              t(case["viewmatrix"][0])[None], t(case["projmatrix"][0])[None], W, H, float(case["tanfov"][0, 0]),
                              float(case["tanfov"][0, 1]), cov3D_precomp=t(cov)[None])
    assert torch.equal(out[0], c_ref)
    # (4) generic channel count (C=3) and no inverse depth requested
    feats3 = np.random.default_rng(0).uniform(size=(P, 3)).astype(np.float32)
    c3 = dict(case); c3["features"] = feats3
    (color3, _, invd3, st3), W, H = mine_forward(c3, 0, render_invdepth=False)
    assert invd3 is None
    assert relerr(color3[0].cpu().numpy(), oracle_forward(c3, 0)["color"]) < TOL
    # (5) markVisible
    rs = R.GaussianRasterizationSettings(H, W, 0.5, 0.5, torch.zeros(3, device=DEV), 1.0, t(case["viewmatrix"][0]), t(case["projmatrix"][0]), 0,
                                         t(case["campos"][0]), False, False, False)
    vis = R.GaussianRasterizer(rs).markVisible(t(np.concatenate([case["means3D"], c2["means3D"][:2]])))
    assert vis[:P].all() and not vis[P:].any()


def test_many_gaussians_and_large_lists_bit_exact():
    """Well beyond the skeletal regime: 600 random Gaussians on a small image (long tile lists, many depth ties impossible
    but thousands of (Gaussian,tile) pairs) -- binning stays bit-identical to the oracle and the image within tolerance."""
    rng = np.random.default_rng(77)
    cfg = small_config(configs.H36M)
    base = raster_case(cfg, seed=2)
    P, C = 600, 3
    case = dict(base)
    case["means3D"] = (rng.uniform(-700, 700, (P, 3)) + np.array([0, 0, 900])).astype(np.float32)
    case["scales"] = np.exp(rng.uniform(2.5, 4.5, (P, 3))).astype(np.float32)
    q = rng.normal(size=(P, 4)).astype(np.float32); case["rotations"] = q / np.linalg.norm(q, axis=1, keepdims=True)
    case["opacities"] = rng.uniform(0.05, 1.0, P).astype(np.float32)
    case["features"] = rng.uniform(size=(P, C)).astype(np.float32)
    (color, radii, invd, st), W, H = mine_forward(case, 0, r_capacity=16384)
    of = oracle_forward(case, 0)
    assert of["R"] > 2000 and st.header()[2] == 0
    assert_stages_equal(st.parse(0), of, radii[0].cpu().numpy(), of["radii"])
    assert relerr(color[0].cpu().numpy(), of["color"]) < TOL
    dL = synthetic_dL(of["color"].shape, 3)
    g = mine_backward(case, 0, st, W, H, dL)
    og = rast.backward(of, case["means3D"], case["scales"], case["rotations"], case["features"], case["viewmatrix"][0], case["projmatrix"][0],
                       W, H, float(case["tanfov"][0, 0]), float(case["tanfov"][0, 1]), dL)
    for k, ok in (("means3D", "dL_dmeans3D"), ("scales", "dL_dscales"), ("rotations", "dL_drotations"), ("features", "dL_dcolors")):
        assert relerr(g[k].reshape(-1), og[ok].reshape(-1)) < 3 * TOL, k


def test_a_million_random_gaussians_bin_bit_exactly():
    """the integer outcome of the projection (radius, tile rectangle => tiles touched, prefix sums, total R) and the
    depth key bits are step functions of fp32 arithmetic whose contraction into FMAs must match the reference expression for
    expression -- checked on >= 10^6 random Gaussians (random anisotropic scales, rotations, opacities, positions incl. behind /
    near the camera plane and off-screen; several random camera rigs of every shape incl. the ragged H36M widths) against the
    UNMODIFIED reference kernels: 0 mismatches in radii, tiles touched, prefix sums and pair count, the fp32 per-Gaussian state
    (depth, 2D mean, conic, cov3D) bit-equal for every visible Gaussian, and -- for the views that fit the op's capacity -- the
    unsorted / sorted keys, values and tile ranges of ~500-Gaussian scenes (thousands of pairs, long tile lists)."""
    from oracle import ref_rasterizer as refr
    if not refr.available("h36m"):
        pytest.skip("oracle/_ref not present on this box")
    P, total, full_checked = 512, 0, 0
    rng = np.random.default_rng(77)
    e = torch.Tensor([]); bg = torch.zeros(32, device=DEV)
    for name, variant, n_seeds, n_frames in (("h36m", "h36m", 5, 40), ("panoptic", "panoptic", 5, 40), ("occlusion-person-8v", "op", 3, 20)):
        cfg = configs.get_config(name)
        C = cfg.n_joints
        feats = np.zeros((P, C), np.float32); feats[np.arange(P), np.arange(P) % C] = 1.0
        for seed in range(n_seeds):
            cams = synthetic.make_cameras(np.random.default_rng(1000 + seed), cfg)
            means = (rng.normal(size=(n_frames, P, 3)) * np.array([900.0, 900.0, 500.0]) + np.array([0.0, 0.0, 900.0])).astype(np.float32)
            means[:, :8] *= 6.0                                         # some far off-screen / behind a camera
            scales = np.exp(rng.uniform(1.0, 4.2, (n_frames, P, 3))).astype(np.float32)
            rots = rng.normal(size=(n_frames, P, 4)).astype(np.float32); rots /= np.linalg.norm(rots, axis=-1, keepdims=True)
            opac = rng.uniform(0.05, 1.0, (n_frames, P)).astype(np.float32)
            for cam in cams:                                            # one batched call per view (own size and field of view)
                W, H, tfx, tfy = cam.image_width, cam.image_height, float(cam.tanfovx), float(cam.tanfovy)
                vmv, pmv = t(cam.world_view_transform), t(cam.full_proj_transform)
                _, radii, _, st = R.rasterize_batched(t(means), t(scales), t(rots), t(opac), t(feats), vmv[None], pmv[None], W, H, tfx, tfy,
                                                      r_capacity=16384, render_invdepth=False)
                radii = radii.cpu().numpy()
                for f in range(n_frames):
                    Rn, _, rradii, geom, binning, img, _ = refr.rasterize_forward(
                        variant, bg, t(means[f]), e, t(opac[f]).reshape(-1, 1), t(scales[f]), t(rots[f]), 1.0, e, vmv, pmv, tfx, tfy, H, W,
                        t(feats).reshape(P, 1, C), 0, t(cam.camera_center), r_capacity=1 << 17)
                    rs = refr.parse_light(refr.RefState(geom, binning, img, Rn, P, W, H, variant))
                    hdr = st.header(f)
                    rr = rradii.cpu().numpy()
                    assert np.array_equal(radii[f], rr)
                    assert int(hdr[7]) == Rn                            # the uncapped pair count
                    if hdr[2] == 0 and f % 4 == 0:                      # fits the capacity: every stage, keys / values / ranges included
                        ms = st.parse(f, W, H)
                        assert_stages_equal(ms, rs, radii[f], rr)
                        full_checked += 1
                    else:
                        fld = lambda fid, dt, n: st._field(f, fid, dt, n)
                        from skelsplat_b200 import lib as L_
                        assert np.array_equal(fld(L_.F_TILES_TOUCHED, np.uint32, P), rs["tiles_touched"])
                        assert np.array_equal(fld(L_.F_POINT_OFFSETS, np.uint32, P), rs["point_offsets"])
                        vis = rr > 0
                        for fid, k, dt, n, shp in ((L_.F_DEPTHS, "depths", np.float32, P, (P,)), (L_.F_MEANS2D, "means2D", np.float32, 2 * P, (P, 2)),
                                                   (L_.F_CONIC_OPACITY, "conic_opacity", np.float32, 4 * P, (P, 4)), (L_.F_COV3D, "cov3D", np.float32, 6 * P, (P, 6))):
                            assert np.array_equal(fld(fid, dt, n).reshape(shp)[vis].view(np.uint32), rs[k][vis].view(np.uint32)), k
                    total += P
    assert total >= 1_000_000 and full_checked >= 100, (total, full_checked)
