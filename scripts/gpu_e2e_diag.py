"""Developer tool (torchrun, N GPUs): where does the multi-GPU e2e step time go?  Times the streaming step with and without the
NCCL gather, with host-side timestamps per call."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import bench
from skelsplat_b200 import configs, trainer

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = f"cuda:{local}"
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(dev))
cfg = configs.H36M; F = 2048
seq, host, gt = bench.make_host_batch(cfg, F, seed=100 + rank)
det = {"poses_2d": torch.from_numpy(host.pop("poses_2d")).pin_memory(), "xyz": torch.from_numpy(host["xyz"]).pin_memory()}
so = trainer.StreamingOptimizer(cfg, seq.cameras, F, int(host["roi_data"].size * 1.1), dev)
gathered = torch.empty((world * F, cfg.n_joints, 3), dtype=torch.float32, device=dev)

side = torch.cuda.Stream(device=dev)

def run(mode, steps=7):
    tickets, marks, evs = [], [], []
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    t_begin = time.perf_counter()
    works = []
    for s in range(steps):
        tickets.append(so.submit_detections(det))
        e1 = torch.cuda.Event(enable_timing=True); e2 = torch.cuda.Event(enable_timing=True)
        e1.record(so.compute_stream)
        src = so.slots[tickets[-1] % 2]["ps"].xyz
        if world > 1 and mode == "gather":
            with torch.cuda.stream(so.compute_stream):
                dist.all_gather_into_tensor(gathered, src)
        elif world > 1 and mode == "gather_side":       # ordered after the optimiser, but the next batch does not wait for it
            side.wait_stream(so.compute_stream)
            with torch.cuda.stream(side):
                dist.all_gather_into_tensor(gathered, src)
        e2.record(so.compute_stream)
        evs.append((e1, e2))
        if len(tickets) >= 2:
            so.result(tickets[-2])
    so.result(tickets[-1]); so.synchronize(); side.synchronize(); torch.cuda.synchronize()
    total = (time.perf_counter() - t_begin) * 1e3
    gather_ms = [round(a.elapsed_time(b), 1) for a, b in evs]
    opt_ms = [round(evs[i][1].elapsed_time(evs[i + 1][0]), 1) for i in range(len(evs) - 1)]
    return {"mode": mode, "ms_per_step": round(total / steps, 1), "gather_ms": gather_ms, "setup+opt_ms": opt_ms}

for mode in ("nogather", "gather", "gather_side", "gather"):
    r = run(mode)
    print(rank, json.dumps(r), flush=True)
if world > 1:
    dist.destroy_process_group()
