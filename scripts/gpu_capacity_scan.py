"""Developer tool: frames that outgrow r_capacity, and speed, per capacity.
usage: gpu_capacity_scan.py <config> <cap,cap,...> [frames]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from skelsplat_b200 import configs, trainer, setup_gpu
name = sys.argv[1] if len(sys.argv) > 1 else "h36m"
cfg = configs.get_config(name)
F = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
CAPS = tuple(int(c) for c in sys.argv[2].split(',')) if len(sys.argv) > 2 else (256, 320, 384, 512)
PCAPS = (0,)
for shard in (0, 3):
    seq, p2d, init0, gt = bench.make_detection_batch(cfg, F, shard)
    ps = setup_gpu.pack_sequence_gpu(cfg, seq.cameras, torch.from_numpy(p2d), torch.from_numpy(init0), "cuda")
    rect = ps.roi_rect.cpu().numpy()
    hw = (rect[..., 2] + rect[..., 3])
    print(json.dumps({"config": name, "shard": shard, "profile_floats_h_plus_w": {"max": int(hw.max()), "p99": float(np.percentile(hw, 99)), "median": float(np.median(hw))}}), flush=True)
    init = tuple(t.clone() for t in (ps.xyz, ps.scaling, ps.rotation, ps.opacity))
    ref = None
    for rcap in CAPS:
        for pc in PCAPS:
            ts = []
            for rep in range(2):
                for d, s_ in zip((ps.xyz, ps.scaling, ps.rotation, ps.opacity), init): d.copy_(s_)
                oc = trainer.make_opt_config(cfg, rcap); lr = trainer.xyz_lr_table(cfg, ps.spatial_lr_scale, oc.iterations)
                loss = torch.empty(F, device="cuda")
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); st = trainer._launch(ps, oc, lr, loss); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ok = (st == 0).cpu().numpy()
            x = ps.xyz.cpu().numpy()
            if ref is None and ok.all():
                ref = x
            same = bool(np.array_equal(x[ok], ref[ok])) if ref is not None else None
            print(json.dumps({"config": name, "shard": shard, "rcap": rcap, "ms": round(min(ts), 2), "fps": round(F / min(ts) * 1e3, 1),
                              "frames_overflowed": int((~ok).sum()), "bit_identical_to_first_complete_run": same}), flush=True)
