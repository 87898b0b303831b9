#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_dropin.py tests/test_gpu_optimizer.py -m gpu -q -x 2>&1 | tail -15
python - <<'PY' 2>&1 | tail -8
import sys, time, torch
sys.path.insert(0, '.')
import bench
from skelsplat_b200 import configs, synthetic
from skelsplat_b200.training import optimise_frame_dropin
cfg = configs.H36M
seq = synthetic.make_sequence(cfg, 3, seed=100)
for graph in (False, True):
    optimise_frame_dropin(seq.frames[0], seq.cameras, cfg, device="cuda", iterations=8, cuda_graph=graph)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    x = optimise_frame_dropin(seq.frames[1], seq.cameras, cfg, device="cuda", iterations=500, cuda_graph=graph)
    torch.cuda.synchronize(); print("graph" if graph else "eager", "frames/s", 1.0 / (time.perf_counter() - t0))
PY
python scripts/gpu_capacity_scan.py h36m 320 2048 2>&1 | tail -2
python scripts/gpu_capacity_scan.py panoptic 1024 2048 2>&1 | tail -1
