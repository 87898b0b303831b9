"""Generates the golden fixtures in this directory FROM THE UNMODIFIED REFERENCE KERNELS.

Run on a GPU box (the reference rasteriser is CUDA-only):
    python tests/golden/make_golden.py [--out tests/golden]
It needs oracle/_ref/libref_rast_*.so (built here from /root/reference by oracle/Makefile; the
.so travels to the GPU box, /root/reference does not).  Fixtures:
  raster_<variant>.npz   seeded small-image inputs + every stage output of the reference forward
                         (R, radii, per-Gaussian state, unsorted/sorted keys and values, ranges, image,
                         inverse depth, final_T, n_contrib) + the reference backward's gradients for a
                         seeded dL (two runs: the reference's atomics make it non-reproducible, the
                         second run records its own spread)
  opt_<config>.npz       final joint positions (500 iterations) of 64 seeded synthetic frames per config from the reference's own
                         Python on the reference's own kernels (see make_opt), plus two re-runs of the first 8 frames (the
                         reference's own spread); configs: h36m, h36m-occ, panoptic, occlusion-person, occlusion-person-8v
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_rasterizer as refr, pipeline as opipe  # noqa: E402
from skelsplat_b200 import configs, synthetic, heatmaps, trainer  # noqa: E402
from skelsplat_b200.cameras import cameras_extent  # noqa: E402
from tests.util import small_config, raster_case, synthetic_dL  # noqa: E402

DEV = "cuda"


def make_raster(variant, cfg_name, out):
    cfg = small_config(configs.get_config(cfg_name))
    case = raster_case(cfg, seed=11)
    e = torch.Tensor([]); bg = torch.zeros(32, device=DEV)
    save = dict(case)
    for vi in range(case["viewmatrix"].shape[0]):
        W, H = int(case["dims"][vi, 0]), int(case["dims"][vi, 1])
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
        means, scales, rots = t(case["means3D"]), t(case["scales"]), t(case["rotations"])
        opac, feats = t(case["opacities"]).reshape(-1, 1), t(case["features"]).reshape(means.shape[0], 1, -1)
        vm, pm, cp = t(case["viewmatrix"][vi]), t(case["projmatrix"][vi]), t(case["campos"][vi])
        tfx, tfy = float(case["tanfov"][vi, 0]), float(case["tanfov"][vi, 1])
        Rn, color, radii, geom, binning, img, invd = refr.rasterize_forward(variant, bg, means, e, opac, scales, rots, 1.0, e, vm, pm,
                                                                            tfx, tfy, H, W, feats, 0, cp)
        st = refr.RefState(geom, binning, img, Rn, means.shape[0], W, H, variant).parse()
        dL = synthetic_dL(tuple(color.shape), seed=vi)          # closed form: recomputed by the tests, not stored
        dLinv = synthetic_dL(tuple(invd.shape), seed=10 + vi)
        grads = []
        for rep in range(2):
            g = refr.rasterize_backward(variant, bg, means, radii, e, opac, scales, rots, 1.0, e, vm, pm, tfx, tfy, t(dL), t(dLinv),
                                        feats, 0, cp, geom, Rn, binning, img)
            grads.append([x.cpu().numpy() for x in g])
        p = f"v{vi}_"
        save.update({p + "R": Rn, p + "radii": radii.cpu().numpy(), p + "color": color.cpu().numpy(), p + "invdepth": invd.cpu().numpy()})
        for k in ("depths", "means2D", "cov3D", "conic_opacity", "tiles_touched", "point_offsets", "point_list", "vals_unsorted",
                  "keys_sorted", "keys_unsorted", "final_T", "n_contrib", "ranges"):
            save[p + k] = st[k]
        names = ("dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales", "dL_drotations")
        for n, a, b in zip(names, grads[0], grads[1]):
            if n == "dL_dsh":
                continue          # garbage in the reference: not part of the contract
            save[p + n] = a
            save[p + n + "_run2"] = b
    np.savez_compressed(os.path.join(out, f"raster_{variant}.npz"), **save)
    print("wrote raster", variant)


def make_opt(cfg_name, out, n_frames=64, seed=1, iterations=500, n_rerun=8):
    """Final poses of the reference pipeline for `n_frames` seeded synthetic frames: the reference's OWN Python (GaussianModel,
    render_*, l2_loss_gaussian, limb_3d_consistency_loss, torch Adam -- imported unmodified, tests/ref_import.py) on the
    reference's OWN kernels (oracle/_ref), under train.py's iteration body (skelsplat_b200/training.py restates it; the loop
    itself needs hydra).  GT heatmaps: the ROI specification (heatmaps.py), the same input our paths consume.  Falls back to
    the restated oracle loop (oracle/pipeline.py, pinned to the reference Python by tests/test_reference_python.py) when the
    staged reference Python is absent.  The first `n_rerun` frames are run two more times: the reference's own spread."""
    from tests import ref_import
    from skelsplat_b200.training import optimise_frame_dropin
    cfg = configs.get_config(cfg_name)
    seq = synthetic.make_sequence(cfg, n_frames, seed=seed)
    ext = cameras_extent(seq.cameras)
    via = "reference-python" if ref_import.available() else "restated-loop"
    if via == "reference-python":
        ref = ref_import.load("ref")
        mods = (ref.gaussian_model.GaussianModel, ref.gaussian_renderer.render_functions, ref.utils.losses, ref.utils.consistency_losses)
    finals, reruns = [], []
    for fi, frame in enumerate(seq.frames):
        _, scal0, rot0, _ = trainer.initial_raw_state(cfg, frame.pose_3d_init[None])
        rois = heatmaps.generate_heatmap_rois(frame.pose_3d_init, frame.poses_2d, seq.cameras, scal0[0], rot0[0])
        dense = [torch.from_numpy(heatmaps.rois_to_dense(rois, v)).to(DEV) for v in range(cfg.nviews)]
        if via == "reference-python":
            run = lambda: optimise_frame_dropin(frame, seq.cameras, cfg, heatmaps_dense=dense, device=DEV, iterations=iterations, modules=mods)
        else:
            run = lambda: opipe.optimise_frame(frame, seq.cameras, cfg, ext, dense, backend="ref", device=DEV, iterations=iterations)
        finals.append(run())
        if fi < n_rerun:        # the reference's backward uses unordered fp32 atomics: two more runs of the SAME frame record its own spread
            reruns.append(np.stack([run() for _ in range(2)]))
    np.savez_compressed(os.path.join(out, f"opt_{cfg_name}.npz"), ref_xyz=np.stack(finals).astype(np.float32), ref_xyz_reruns=np.stack(reruns).astype(np.float32),
                        seed=seed, n_frames=n_frames, iterations=iterations, n_rerun=n_rerun, via=via,
                        init_xyz=np.stack([f.pose_3d_init for f in seq.frames]), gt_xyz=np.stack([f.pose_3d_gt for f in seq.frames]))
    mine = trainer.optimize_sequence(seq, DEV, iterations=iterations)        # calibration printout only (not stored)
    ref_xyz, rr = np.stack(finals), np.stack(reruns)
    dev = np.linalg.norm(mine - ref_xyz, axis=-1)
    spread = np.linalg.norm(rr - ref_xyz[:n_rerun, None], axis=-1)
    print("wrote opt", cfg_name, via, "| fused vs golden: max %.4f p99 %.4f median %.5f mm | reference spread: max %.4f p99 %.4f median %.5f mm | mpjpe delta %.5f"
          % (dev.max(), np.percentile(dev, 99), np.median(dev), spread.max(), np.percentile(spread, 99), np.median(spread),
             trainer.mpjpe(mine, np.stack([f.pose_3d_gt for f in seq.frames])) - trainer.mpjpe(ref_xyz, np.stack([f.pose_3d_gt for f in seq.frames]))), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "golden"))
    ap.add_argument("--frames", type=int, default=64, help="frames per config (SURVEY.md 8d: parity sets of 64)")
    ap.add_argument("--skip-raster", action="store_true")
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    if not args.skip_raster:
        for variant, name in (("h36m", "h36m"), ("panoptic", "panoptic"), ("op", "occlusion-person")):
            make_raster(variant, name, args.out)
    for name in ("h36m", "h36m-occ", "panoptic", "occlusion-person", "occlusion-person-8v"):      # 8v: the stale/zero-slot accumulation quirk at 500 iterations
        make_opt(name, args.out, n_frames=args.frames)
