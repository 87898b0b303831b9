"""Developer tool: frames/s + checksum of the fused optimiser on the bench workload (bitwise regression check between kernel versions)."""
import os, sys, json, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
from skelsplat_b200 import configs, trainer
for name, F in (("h36m", 2048), ("occlusion-person-8v", 2048), ("panoptic", 1024)):
    cfg = configs.get_config(name)
    seq, host, gt = bench.make_host_batch(cfg, F, seed=100)
    ps = trainer.pack_sequence(cfg, seq.cameras, host["xyz"], None, "cuda", host=host)
    init = tuple(t.clone() for t in (ps.xyz, ps.scaling, ps.rotation, ps.opacity))
    ts = []
    for rep in range(3):
        for d, s_ in zip((ps.xyz, ps.scaling, ps.rotation, ps.opacity), init): d.copy_(s_)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); trainer.optimize_packed(ps, check=(rep == 0)); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    x = ps.xyz.cpu().numpy()
    print(json.dumps({"config": name, "frames": F, "ms": round(min(ts), 2), "fps": round(F / min(ts) * 1e3, 1),
                      "checksum": float(np.abs(x.astype(np.float64)).sum()), "mpjpe": trainer.mpjpe(x, gt)}), flush=True)
