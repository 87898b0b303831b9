"""Developer tool: a small pass through every kernel of the library, for compute-sanitizer (memcheck / racecheck).
   compute-sanitizer --tool memcheck python scripts/gpu_sanitize_target.py [iterations | ssim]     (ssim: only the loss / SSIM kernels)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from skelsplat_b200 import configs, synthetic, trainer, setup_gpu, loss_utils, rasterizer as R
from tests.util import small_config, raster_case

dev = "cuda"
only_ssim = len(sys.argv) > 1 and sys.argv[1] == "ssim"
it = int(sys.argv[1]) if len(sys.argv) > 1 and not only_ssim else 8
for base in (() if only_ssim else (configs.H36M, configs.PANOPTIC, configs.OCCLUSION_PERSON_8V)):
    cfg = small_config(base, 4)
    seq = synthetic.make_sequence(cfg, 2, seed=3)
    p2 = np.stack([f.poses_2d for f in seq.frames]).astype(np.float32)
    # fused optimiser through the host-prepared and the GPU-prepared paths, and the streaming API
    print(cfg.name, "optimise", trainer.optimize_sequence(seq, dev, iterations=it).shape, flush=True)
    ps = setup_gpu.pack_sequence_gpu(cfg, seq.cameras, p2, None, dev)
    trainer.optimize_packed(ps, iterations=it)
    so = trainer.StreamingOptimizer(cfg, seq.cameras, 2, int(ps.roi_data.numel()) + 64, dev, iterations=it)
    so.result(so.submit_detections({"poses_2d": torch.from_numpy(p2).pin_memory()}))
    # dense-contract rasteriser forward + backward, two views of one size, non one-hot features
    case = raster_case(cfg, seed=1, n_views=2)
    t = {k: torch.from_numpy(v).to(dev) for k, v in case.items()}
    W, H = int(case["dims"][:, 0].max()), int(case["dims"][:, 1].max())
    tfx, tfy = float(case["tanfov"][0, 0]), float(case["tanfov"][0, 1])
    args = (t["means3D"][None], t["scales"][None], t["rotations"][None], t["opacities"][None], t["features"],
            t["viewmatrix"], t["projmatrix"], W, H, tfx, tfy)
    color, radii, invd, st = R.rasterize_batched(*args)
    R.rasterize_batched_backward(st, *args, torch.ones_like(color), torch.ones_like(invd))
    torch.cuda.synchronize()
    print(cfg.name, "dense forward/backward ok", flush=True)
# dense losses + SSIM
r = torch.rand(5, 70, 90, device=dev, requires_grad=True); g = torch.rand(5, 70, 90, device=dev) * (torch.rand(5, 70, 90, device=dev) > 0.5)
l, _ = loss_utils.l2_loss_gaussian(r, g, None); l.backward()
from fused_ssim import fused_ssim
a = torch.rand(2, 3, 50, 70, device=dev, requires_grad=True); b = torch.rand(2, 3, 50, 70, device=dev)
fused_ssim(a, b).backward()
# the map-form SSIM (FusedSSIMMap: the reference's autograd surface) incl. 'valid' padding
import fused_ssim as FS
m = FS.FusedSSIMMap.apply(0.01 ** 2, 0.03 ** 2, a, b, "valid", True); m.sum().backward()
for shape in ((1, 1, 13, 9), (1, 2, 97, 131), (2, 1, 33, 300)):      # ragged edges: partial strips, W < 32, H < one strip
    a = torch.rand(*shape, device=dev, requires_grad=True); b = torch.rand(*shape, device=dev)
    fused_ssim(a, b).backward(); fused_ssim(a.detach(), b, train=False)
    FS.FusedSSIMMap.apply(0.01 ** 2, 0.03 ** 2, a, b, "same", True).sum().backward()
if only_ssim:
    torch.cuda.synchronize(); print("SANITIZE TARGET DONE (ssim only)"); sys.exit(0)
# the dense loop's iteration bodies + the graph-capturable Adam kernel, launched eagerly (no capture under the sanitizer)
from skelsplat_b200.training import GraphedFrameOptimizer
cfg = small_config(configs.OCCLUSION_PERSON_8V, 4)
seq = synthetic.make_sequence(cfg, 1, seed=5)
gfo = GraphedFrameOptimizer(cfg, seq.cameras, dev, iterations=16)
gfo._reset(seq.frames[0].pose_3d_init)
_, scal0, rot0, _ = trainer.initial_raw_state(cfg, seq.frames[0].pose_3d_init[None])
from skelsplat_b200 import heatmaps
gfo.load_heatmaps(rois=heatmaps.generate_heatmap_rois(seq.frames[0].pose_3d_init, seq.frames[0].poses_2d, seq.cameras, scal0[0], rot0[0]))
for pat in range(gfo.n_patterns):
    gfo._step_group(pat)
# the debug accessor of the fused kernel's binning state
ps = setup_gpu.pack_sequence_gpu(cfg, seq.cameras, np.stack([f.poses_2d for f in seq.frames]).astype(np.float32), None, dev)
trainer.debug_binning(ps, frame=0, step=1)
torch.cuda.synchronize()
print("SANITIZE TARGET DONE")
