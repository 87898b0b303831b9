"""Developer tool: frames/s + checksum of the fused optimiser on the bench workloads (bitwise regression check between kernel
versions / launch shapes).  Inputs come from the GPU setup path (detections -> DLT init -> heatmap ROIs), like bench.py's.
For every config the launch shapes selectable through ssb_opt_config::resident_record_slots are run on the SAME inputs and
their results compared bit for bit (Panoptic: one 1024-thread CTA/SM with all records resident vs two 512-thread CTAs/SM with
the records of two slots at a time)."""
import os, sys, json, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
from skelsplat_b200 import configs, trainer, setup_gpu
for name, F in (("h36m", 2048), ("h36m-occ", 2048), ("occlusion-person-8v", 2048), ("panoptic", 2048)):
    cfg = configs.get_config(name)
    seq, p2d, init, gt = bench.make_detection_batch(cfg, F, 0)
    ps = setup_gpu.pack_sequence_gpu(cfg, seq.cameras, torch.from_numpy(p2d), torch.from_numpy(init), "cuda")
    init_state = tuple(t.clone() for t in (ps.xyz, ps.scaling, ps.rotation, ps.opacity))
    results = {}
    for mode in ("0", "4", "2"):                  # auto | all four slots' records resident | two at a time
        os.environ["SKELSPLAT_B200_RECORD_SLOTS"] = mode
        ts, over = [], 0
        try:
            for rep in range(3):
                for d, s_ in zip((ps.xyz, ps.scaling, ps.rotation, ps.opacity), init_state): d.copy_(s_)
                oc = trainer.make_opt_config(cfg, trainer.default_r_capacity(cfg))
                lr = trainer.xyz_lr_table(cfg, ps.spatial_lr_scale, oc.iterations)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); st = trainer._launch(ps, oc, lr, torch.empty(F, device="cuda")); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1)); over = int((st != 0).sum())
        except Exception as e:      # a forced shape that does not fit in shared memory
            print(json.dumps({"config": name, "record_slots": mode, "error": repr(e)[:200]}), flush=True)
            continue
        x = ps.xyz.cpu().numpy()
        results[mode] = x
        print(json.dumps({"config": name, "record_slots": mode, "frames": F, "r_capacity": trainer.default_r_capacity(cfg), "ms": round(min(ts), 2),
                          "fps": round(F / min(ts) * 1e3, 1), "frames_over_capacity": over,
                          "checksum": float(np.abs(x.astype(np.float64)).sum()), "mpjpe": trainer.mpjpe(x, gt)}), flush=True)
    ks = list(results)
    print(json.dumps({"config": name, "bit_identical_across_launch_shapes": all(np.array_equal(results[ks[0]], results[k]) for k in ks[1:]),
                      "shapes": ks}), flush=True)
os.environ.pop("SKELSPLAT_B200_RECORD_SLOTS", None)
