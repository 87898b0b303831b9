"""Developer tool: frames that outgrow r_capacity, and speed, per capacity (H36M bench shape, the 8 rank seeds)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from skelsplat_b200 import configs, trainer
cfg = configs.get_config(sys.argv[1]) if len(sys.argv) > 1 else configs.H36M
F = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
CAPS = tuple(int(c) for c in sys.argv[2].split(',')) if len(sys.argv) > 2 else (256, 320, 384, 512)
for seed in (100, 103, 106):
    seq, host, gt = bench.make_host_batch(cfg, F, seed=seed)
    host.pop("poses_2d")
    ps = trainer.pack_sequence(cfg, seq.cameras, host["xyz"], None, "cuda", host=host)
    init = tuple(t.clone() for t in (ps.xyz, ps.scaling, ps.rotation, ps.opacity))
    for rcap in CAPS:
        ts = []
        for rep in range(2):
            for d, s_ in zip((ps.xyz, ps.scaling, ps.rotation, ps.opacity), init): d.copy_(s_)
            oc = trainer.make_opt_config(cfg, rcap); lr = trainer.xyz_lr_table(cfg, ps.spatial_lr_scale, oc.iterations)
            loss = torch.empty(F, device="cuda")
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); st = trainer._launch(ps, oc, lr, loss); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print(json.dumps({"seed": seed, "rcap": rcap, "ms": round(min(ts), 2), "fps": round(F / min(ts) * 1e3, 1),
                          "frames_overflowed": int((st != 0).sum())}), flush=True)
