"""Small workloads for ncu captures (developer tool).  usage: profile_target.py opt|dense|loss [frames]"""
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
    seq, host, gt = bench.make_host_batch(cfg, F, seed=100)
    ps = trainer.pack_sequence(cfg, seq.cameras, host["xyz"], None, dev, host=host)
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
