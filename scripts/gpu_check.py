"""GPU-box bring-up script (developer tool, not part of the product or the test-suite).

Runs every CUDA path once against the UNMODIFIED reference kernels (oracle/_ref) and the C oracle,
prints a compact report and writes gpurun_out/gpu_check.json (+ golden fixtures under
gpurun_out/golden/ via tests/golden/make_golden.py when --golden is given).
"""
import argparse
import json
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from skelsplat_b200 import configs, synthetic, heatmaps, trainer  # noqa: E402
from skelsplat_b200 import rasterizer as R  # noqa: E402
from skelsplat_b200.cameras import cameras_extent  # noqa: E402
from oracle import rast as crast, ref_rasterizer as refr, pipeline as opipe  # noqa: E402

REPORT = {}
DEV = "cuda"


def section(name):
    def deco(fn):
        def run(*a, **k):
            t = time.time()
            try:
                REPORT[name] = fn(*a, **k)
                REPORT[name]["_seconds"] = round(time.time() - t, 2)
                print(f"[{name}] {json.dumps(REPORT[name])}", flush=True)
            except Exception as e:  # keep going: one broken section must not hide the others
                REPORT[name] = {"error": repr(e), "trace": traceback.format_exc()[-1500:]}
                print(f"[{name}] FAILED {e}\n{traceback.format_exc()}", flush=True)
        return run
    return deco


def variant_of(cfg):
    return opipe.VARIANT_OF[cfg.rendering]


def gaussians_for(cfg, frame, device=DEV):
    J = cfg.n_joints
    xyz, scal, rot, opa = trainer.initial_raw_state(cfg, frame.pose_3d_init[None])
    t = lambda a: torch.from_numpy(a).to(device)
    means = t(xyz[0]); scales = torch.exp(t(scal[0])); rots = torch.nn.functional.normalize(t(rot[0]))
    opac = torch.sigmoid(t(opa[0])).reshape(J, 1)
    feats = torch.eye(J, device=device).reshape(J, 1, J).contiguous()
    return means, scales, rots, opac, feats


def relerr(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    d = np.abs(a - b).max() if a.size else 0.0
    s = np.abs(b).max() if b.size else 0.0
    return float(d / s) if s > 0 else float(d)


def ref_forward(cfg, cam, means, scales, rots, opac, feats):
    v = variant_of(cfg)
    bg = torch.zeros(32, device=DEV)
    e = torch.Tensor([])
    vm = torch.from_numpy(cam.world_view_transform).to(DEV); pm = torch.from_numpy(cam.full_proj_transform).to(DEV)
    cp = torch.from_numpy(cam.camera_center).to(DEV)
    out = refr.rasterize_forward(v, bg, means, e, opac, scales, rots, 1.0, e, vm, pm, cam.tanfovx, cam.tanfovy,
                                 cam.image_height, cam.image_width, feats, 0, cp)
    return out, (bg, vm, pm, cp)


def ref_backward(cfg, cam, means, scales, rots, opac, feats, fwd_out, aux, dL, dLinv):
    v = variant_of(cfg)
    Rn, color, radii, geom, binning, img, invd = fwd_out
    bg, vm, pm, cp = aux
    e = torch.Tensor([])
    return refr.rasterize_backward(v, bg, means, radii, e, opac, scales, rots, 1.0, e, vm, pm, cam.tanfovx, cam.tanfovy,
                                   dL, dLinv, feats, 0, cp, geom, Rn, binning, img)


@section("dense_parity")
def dense_parity():
    res = {}
    for name in ("h36m", "panoptic", "occlusion-person"):
        cfg = configs.get_config(name)
        seq = synthetic.make_sequence(cfg, 2, seed=0)
        worst = dict(bits=0, color=0.0, invdepth=0.0, grads_vs_ref={}, ref_spread={}, grads_vs_oracle={}, color_vs_oracle=0.0)
        for fi, frame in enumerate(seq.frames):
            means, scales, rots, opac, feats = gaussians_for(cfg, frame)
            # perturb scales/rotations so that the anisotropic / rotated paths are exercised
            g = torch.Generator(device="cpu").manual_seed(fi)
            scales = scales * torch.exp(0.3 * torch.randn(scales.shape, generator=g)).to(DEV)
            rots = torch.nn.functional.normalize(rots + 0.3 * torch.randn(rots.shape, generator=g).to(DEV))
            for cam in seq.cameras:
                W, H = cam.image_width, cam.image_height
                fo, aux = ref_forward(cfg, cam, means, scales, rots, opac, feats)
                Rn, rcolor, rradii, geom, binning, img, rinvd = fo
                rs = refr.RefState(geom, binning, img, Rn, cfg.n_joints, W, H, variant_of(cfg)).parse()
                vm = aux[1].reshape(1, 4, 4); pm = aux[2].reshape(1, 4, 4)
                color, radii, invd, st = R.rasterize_batched(means[None], scales[None], rots[None], opac.reshape(1, -1),
                                                             feats.reshape(cfg.n_joints, -1), vm, pm, W, H, cam.tanfovx, cam.tanfovy)
                torch.cuda.synchronize()
                ms = st.parse(0)
                vis = rradii.cpu().numpy() > 0
                bits = 0
                bits += int(ms["R"] != Rn)
                bits += int(not np.array_equal(radii[0].cpu().numpy(), rradii.cpu().numpy()))
                for k in ("keys_unsorted", "vals_unsorted", "keys_sorted", "point_list", "ranges", "tiles_touched", "point_offsets"):
                    bits += int(not np.array_equal(ms[k], rs[k]))
                for k in ("depths", "means2D", "conic_opacity", "cov3D"):
                    bits += int(not np.array_equal(ms[k][vis].view(np.uint32), rs[k][vis].view(np.uint32)))
                worst["bits"] += bits
                worst["color"] = max(worst["color"], relerr(color[0].cpu().numpy(), rcolor.cpu().numpy()))
                worst["invdepth"] = max(worst["invdepth"], relerr(invd[0].cpu().numpy(), rinvd.cpu().numpy()))
                # backward with the l2_gaussian gradient of a shifted GT (so gradients are non-trivial)
                gt = torch.roll(rcolor, shifts=(2, -3), dims=(1, 2)) * 0.8
                mask = (gt > 0) | (rcolor > 0)
                dL = torch.where(mask, 2 * (rcolor - gt) / mask.sum(), torch.zeros_like(rcolor)).contiguous()
                dLinv = torch.zeros_like(rinvd)
                rb1 = ref_backward(cfg, cam, means, scales, rots, opac, feats, fo, aux, dL, dLinv)
                rb2 = ref_backward(cfg, cam, means, scales, rots, opac, feats, fo, aux, dL, dLinv)
                mb = R.rasterize_batched_backward(st, means[None], scales[None], rots[None], opac.reshape(1, -1),
                                                  feats.reshape(cfg.n_joints, -1), vm, pm, W, H, cam.tanfovx, cam.tanfovy, dL[None], dLinv[None])
                torch.cuda.synchronize()
                names = dict(means2D=0, features=1, opacity=2, means3D=3, cov3D=4, scales=6, rotations=7)
                for k, idx in names.items():
                    a = mb[k][0].cpu().numpy().reshape(-1); b = rb1[idx].cpu().numpy().reshape(-1); c = rb2[idx].cpu().numpy().reshape(-1)
                    worst["grads_vs_ref"][k] = max(worst["grads_vs_ref"].get(k, 0.0), relerr(a, b))
                    worst["ref_spread"][k] = max(worst["ref_spread"].get(k, 0.0), relerr(c, b))
                if fi == 0 and cam.uid == 0:
                    of = crast.forward(means.cpu().numpy(), scales.cpu().numpy(), rots.cpu().numpy(), opac.cpu().numpy(),
                                       feats.reshape(cfg.n_joints, -1).cpu().numpy(), cam.world_view_transform, cam.full_proj_transform,
                                       W, H, cam.tanfovx, cam.tanfovy)
                    ob = crast.backward(of, means.cpu().numpy(), scales.cpu().numpy(), rots.cpu().numpy(),
                                        feats.reshape(cfg.n_joints, -1).cpu().numpy(), cam.world_view_transform, cam.full_proj_transform,
                                        W, H, cam.tanfovx, cam.tanfovy, dL.cpu().numpy(), dLinv.cpu().numpy())
                    worst["oracle_bits"] = int(not np.array_equal(of["keys_sorted"], rs["keys_sorted"])) + int(not np.array_equal(of["ranges"], rs["ranges"])) + int(not np.array_equal(of["point_list"], rs["point_list"]))
                    worst["color_vs_oracle"] = relerr(color[0].cpu().numpy(), of["color"])
                    for k, ok in dict(means3D="dL_dmeans3D", scales="dL_dscales", rotations="dL_drotations", means2D="dL_dmeans2D").items():
                        worst["grads_vs_oracle"][k] = relerr(mb[k][0].cpu().numpy().reshape(-1), ob[ok].reshape(-1))
        res[name] = worst
    return res


@section("stress_bits")
def stress_bits(n_views=200, P=256):
    """Random Gaussians: every per-Gaussian output and every key must equal the reference bit for bit."""
    cfg = configs.H36M
    rng = np.random.default_rng(123)
    seq = synthetic.make_sequence(cfg, 1, seed=5)
    mism = dict(radii=0, keys=0, sorted=0, ranges=0, state=0, R=0)
    total = 0
    e = torch.Tensor([])
    bg = torch.zeros(32, device=DEV)
    for it in range(n_views):
        cam = seq.cameras[it % 4]
        W, H = cam.image_width, cam.image_height
        means = torch.from_numpy((rng.uniform(-1500, 1500, (P, 3)) + np.array([0, 0, 900])).astype(np.float32)).to(DEV)
        scales = torch.from_numpy(np.exp(rng.uniform(0.5, 4.5, (P, 3))).astype(np.float32)).to(DEV)
        rots = torch.from_numpy(rng.normal(size=(P, 4)).astype(np.float32)).to(DEV)
        rots = torch.nn.functional.normalize(rots)
        opac = torch.from_numpy(rng.uniform(0.05, 1.0, (P, 1)).astype(np.float32)).to(DEV)
        feats = torch.from_numpy(rng.uniform(0, 1, (P, 1, 17)).astype(np.float32)).to(DEV)
        vm = torch.from_numpy(cam.world_view_transform).to(DEV); pm = torch.from_numpy(cam.full_proj_transform).to(DEV)
        cp = torch.from_numpy(cam.camera_center).to(DEV)
        Rn, rcolor, rradii, geom, binning, img, rinvd = refr.rasterize_forward("h36m", bg, means, e, opac, scales, rots, 1.0, e, vm, pm,
                                                                               cam.tanfovx, cam.tanfovy, H, W, feats, 0, cp, r_capacity=1 << 16)
        rs = refr.RefState(geom, binning, img, Rn, P, W, H, "h36m").parse()
        color, radii, invd, st = R.rasterize_batched(means[None], scales[None], rots[None], opac.reshape(1, -1), feats.reshape(P, 17),
                                                     vm.reshape(1, 4, 4), pm.reshape(1, 4, 4), W, H, cam.tanfovx, cam.tanfovy, r_capacity=8192)
        torch.cuda.synchronize()
        ms = st.parse(0)
        vis = rradii.cpu().numpy() > 0
        total += P
        mism["R"] += int(ms["R"] != Rn)
        mism["radii"] += int((radii[0].cpu().numpy() != rradii.cpu().numpy()).sum())
        if ms["R"] == Rn:
            mism["keys"] += int((ms["keys_unsorted"] != rs["keys_unsorted"]).sum())
            mism["sorted"] += int((ms["keys_sorted"] != rs["keys_sorted"]).sum() + (ms["point_list"] != rs["point_list"]).sum())
            mism["ranges"] += int((ms["ranges"] != rs["ranges"]).sum())
        for k in ("depths", "means2D", "conic_opacity", "cov3D"):
            mism["state"] += int((ms[k][vis].view(np.uint32) != rs[k][vis].view(np.uint32)).sum())
        if it == 0:
            mism["color_rel"] = relerr(color[0].cpu().numpy(), rcolor.cpu().numpy())
            mism["R_example"] = Rn
    mism["gaussians"] = total
    return mism


@section("losses")
def losses():
    from skelsplat_b200 import loss_utils as LU
    out = {}
    torch.manual_seed(0)
    r = torch.rand(17, 250, 333, device=DEV) * (torch.rand(17, 250, 333, device=DEV) > 0.7)
    g = torch.rand(17, 250, 333, device=DEV) * (torch.rand(17, 250, 333, device=DEV) > 0.8)
    for name, mine, ref in (("l2_gaussian", lambda a, b: LU.l2_loss_gaussian(a, b, None)[0], lambda a, b: opipe.l2_loss_gaussian(a, b)[0]),
                            ("l1", lambda a, b: LU.l1_loss(a, b, None), opipe.l1_loss),
                            ("l1_gaussian", lambda a, b: LU.l1_loss_gaussian(a, b, None), opipe.l1_loss_gaussian)):
        a = r.clone().requires_grad_(True); b = r.clone().requires_grad_(True)
        lm = mine(a, g); lr_ = ref(b, g)
        lm.backward(); lr_.backward()
        out[name] = dict(value=relerr(lm.item(), lr_.item()), grad=relerr(a.grad.cpu().numpy(), b.grad.cpu().numpy()))
    xyz = torch.randn(5, 17, 3, device=DEV) * 300
    pairs = configs.H36M.limb_pairs
    a = xyz.clone().requires_grad_(True)
    ref = torch.stack([opipe.limb_3d_consistency_loss(a[i], pairs) for i in range(5)])
    ref.sum().backward()
    b = xyz.clone().requires_grad_(True)
    mine = LU.limb_3d_consistency_loss_batched(b, pairs)
    mine.sum().backward()
    out["limb"] = dict(value=relerr(mine.detach().cpu().numpy(), ref.detach().cpu().numpy()), grad=relerr(b.grad.cpu().numpy(), a.grad.cpu().numpy()))
    return out


@section("ssim")
def ssim():
    from fused_ssim import fused_ssim
    torch.manual_seed(0)
    out = {}
    for shape in ((2, 3, 120, 200), (1, 5, 97, 61)):
        img1 = torch.rand(shape, device=DEV, requires_grad=True)
        img2 = torch.rand(shape, device=DEV)
        a = fused_ssim(img1, img2, "same")
        a.backward()
        ga = img1.grad.clone(); img1.grad = None
        b = opipe.ssim(img1, img2)
        b.backward()
        out[str(shape)] = dict(value=relerr(a.item(), b.item()), grad=relerr(ga.cpu().numpy(), img1.grad.cpu().numpy()),
                               isclose=bool(torch.isclose(a, b)), grad_isclose=bool(torch.isclose(ga, img1.grad, rtol=1e-5, atol=1e-7).all()))
    return out


@section("optimiser_parity")
def optimiser_parity(n_frames=4, iterations=500):
    res = {}
    for name in ("h36m", "h36m-occ", "panoptic", "occlusion-person"):
        cfg = configs.get_config(name)
        seq = synthetic.make_sequence(cfg, n_frames, seed=1)
        ext = cameras_extent(seq.cameras)
        t0 = time.time()
        mine = trainer.optimize_sequence(seq, DEV, iterations=iterations)
        torch.cuda.synchronize()
        t_mine = time.time() - t0
        refs, refs2 = [], []
        t_ref = 0.0
        for fi, frame in enumerate(seq.frames):
            J = cfg.n_joints
            xyz0, scal0, rot0, _ = trainer.initial_raw_state(cfg, frame.pose_3d_init[None])
            rois = heatmaps.generate_heatmap_rois(frame.pose_3d_init, frame.poses_2d, seq.cameras, scal0[0], rot0[0])
            dense = [torch.from_numpy(heatmaps.rois_to_dense(rois, v)).to(DEV) for v in range(cfg.nviews)]
            torch.cuda.synchronize(); t1 = time.time()
            refs.append(opipe.optimise_frame(frame, seq.cameras, cfg, ext, dense, backend="ref", device=DEV, iterations=iterations))
            torch.cuda.synchronize(); t_ref += time.time() - t1
            if fi == 0:
                refs2.append(opipe.optimise_frame(frame, seq.cameras, cfg, ext, dense, backend="ref", device=DEV, iterations=iterations))
        refs = np.stack(refs)
        gt = np.stack([f.pose_3d_gt for f in seq.frames]); init = np.stack([f.pose_3d_init for f in seq.frames])
        dev = np.linalg.norm(mine - refs, axis=-1)
        res[name] = dict(max_joint_dev_mm=float(dev.max()), mean_joint_dev_mm=float(dev.mean()),
                         ref_spread_mm=float(np.linalg.norm(refs2[0] - refs[0], axis=-1).max()),
                         mpjpe_init=trainer.mpjpe(init, gt), mpjpe_ref=trainer.mpjpe(refs, gt), mpjpe_mine=trainer.mpjpe(mine, gt),
                         ref_s_per_frame=t_ref / n_frames, mine_s_total=t_mine)
        os.makedirs(os.path.join(ROOT, "gpurun_out", "golden"), exist_ok=True)
        np.savez_compressed(os.path.join(ROOT, "gpurun_out", "golden", f"opt_{name}.npz"), ref_xyz=refs, mine_xyz=mine,
                            seed=1, n_frames=n_frames, iterations=iterations)
    return res


@section("throughput")
def throughput(F=2048):
    res = {}
    for name in ("h36m", "occlusion-person-8v", "panoptic"):
        cfg = configs.get_config(name)
        seq = synthetic.make_sequence(cfg, 64, seed=2)
        poses_init = np.stack([f.pose_3d_init for f in seq.frames]); poses_2d = np.stack([f.poses_2d for f in seq.frames])
        host = trainer.pack_host(cfg, seq.cameras, poses_init, poses_2d)
        reps = F // 64
        big = dict(host)
        for k in ("xyz", "scaling", "rotation", "opacity", "roi_rect"):
            big[k] = np.concatenate([host[k]] * reps)
        big["roi_offset"] = np.concatenate([host["roi_offset"]] * reps)   # frames share ROI data (L2-friendly; noted)
        ps = trainer.pack_sequence(cfg, seq.cameras, np.concatenate([poses_init] * reps), None, DEV, host=big)
        state0 = (ps.xyz.clone(), ps.scaling.clone(), ps.rotation.clone(), ps.opacity.clone())
        times = []
        for rep in range(3):
            ps.xyz.copy_(state0[0]); ps.scaling.copy_(state0[1]); ps.rotation.copy_(state0[2]); ps.opacity.copy_(state0[3])
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            trainer.optimize_packed(ps, check=False)
            e1.record(); torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1) / 1e3)
        res[name] = dict(frames=F, seconds=min(times), frames_per_s=F / min(times))
    return res


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--sections", default="dense_parity,stress_bits,losses,ssim,optimiser_parity,throughput")
    ap.add_argument("--opt-frames", type=int, default=4)
    ap.add_argument("--opt-iterations", type=int, default=500)
    args = ap.parse_args()
    print(torch.cuda.get_device_name(0), flush=True)
    todo = args.sections.split(",")
    if "dense_parity" in todo: dense_parity()
    if "stress_bits" in todo: stress_bits()
    if "losses" in todo: losses()
    if "ssim" in todo: ssim()
    if "optimiser_parity" in todo: optimiser_parity(args.opt_frames, args.opt_iterations)
    if "throughput" in todo: throughput()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "gpu_check.json"), "w") as f:
        json.dump(REPORT, f, indent=1)
