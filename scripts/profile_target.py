"""Small workloads for ncu captures (developer tool).  usage: profile_target.py opt|dense|loss|ssim [frames] [config]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from skelsplat_b200 import configs, synthetic, trainer
from skelsplat_b200 import rasterizer as R
import bench

what = sys.argv[1] if len(sys.argv) > 1 else "opt"
dev = "cuda"
cfg = configs.get_config(sys.argv[3]) if len(sys.argv) > 3 else configs.H36M
if what == "opt":
    F = int(sys.argv[2]) if len(sys.argv) > 2 else 592
    from skelsplat_b200 import setup_gpu
    seq, p2d, init, gt = bench.make_detection_batch(cfg, F, 0)
    ps = setup_gpu.pack_sequence_gpu(cfg, seq.cameras, torch.from_numpy(p2d), torch.from_numpy(init), dev)
    trainer.optimize_packed(ps, check=False)
    torch.cuda.synchronize()
elif what == "dense":
    seq = synthetic.make_sequence(cfg, 8, seed=100)
    out = bench.bench_dense_rasterizer(torch, R, cfg, seq, dev, 6550.7, frames=4, reps=1)
    print(out)
elif what == "loss":
    from skelsplat_b200 import loss_utils as LU
    r = torch.rand(17, 1000, 1000, device=dev, requires_grad=True); g = torch.rand(17, 1000, 1000, device=dev)
    for _ in range(3):
        l, _ = LU.l2_loss_gaussian(r, g, None, want_error=False); l.backward()
    torch.cuda.synchronize()
elif what == "ssim":
    from fused_ssim import fused_ssim
    a = torch.rand(5, 1, 1500, 1500, device=dev).requires_grad_(True); b = torch.rand(5, 1, 1500, 1500, device=dev)
    for _ in range(2):
        a.grad = None
        fused_ssim(a, b).backward()
        fused_ssim(a.detach(), b, train=False)
    torch.cuda.synchronize()
elif what == "setup":
    from skelsplat_b200 import setup_gpu
    seq, p2d, init, gt = bench.make_detection_batch(cfg, 2048, 0)
    for _ in range(2):
        ps = setup_gpu.pack_sequence_gpu(cfg, seq.cameras, torch.from_numpy(p2d), None, dev)      # DLT + ROI rects / offsets / profiles
    torch.cuda.synchronize()
