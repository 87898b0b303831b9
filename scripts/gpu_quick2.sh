#!/bin/bash
mkdir -p gpurun_out
echo "--- adam form 1 (product library): fma(value, g*g, self)"
python -m pytest tests/test_gpu_losses.py -m gpu -q -k adam 2>&1 | tail -4
echo "--- adam form 0: fma(value*g, g, self)"
SKELSPLAT_B200_LIB=skelsplat_b200/_lib/libvariant_adam0.so python -m pytest tests/test_gpu_losses.py -m gpu -q -k adam 2>&1 | tail -4
python -m pytest tests/test_gpu_dropin.py -m gpu -q 2>&1 | tail -6
python - <<'PY' 2>&1 | tail -6
import sys, time, torch, numpy as np
sys.path.insert(0, '.')
from skelsplat_b200 import configs, synthetic, trainer, heatmaps
from skelsplat_b200.training import GraphedFrameOptimizer
cfg = configs.H36M
seq = synthetic.make_sequence(cfg, 6, seed=100)
t0 = time.perf_counter(); gfo = GraphedFrameOptimizer(cfg, seq.cameras, "cuda"); gfo.capture(); torch.cuda.synchronize(); print("construct+capture %.0f ms" % ((time.perf_counter() - t0) * 1e3))
rois = []
for fr in seq.frames:
    _, scal0, rot0, _ = trainer.initial_raw_state(cfg, fr.pose_3d_init[None])
    rois.append(heatmaps.generate_heatmap_rois(fr.pose_3d_init, fr.poses_2d, seq.cameras, scal0[0], rot0[0]))
gfo.optimise(seq.frames[0].pose_3d_init, rois=rois[0])
torch.cuda.synchronize(); t0 = time.perf_counter()
for fr, r in zip(seq.frames[1:], rois[1:]):
    x = gfo.optimise(fr.pose_3d_init, rois=r)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print("graphed dense loop: %.1f ms/frame = %.2f frames/s" % (dt * 1e3, 1 / dt))
fused = trainer.optimize_sequence(synthetic.Sequence(cfg=cfg, cameras=seq.cameras, frames=seq.frames[5:6]), "cuda")[0]
print("vs fused optimiser max joint deviation (mm):", float(np.linalg.norm(x - fused, axis=-1).max()))
PY
