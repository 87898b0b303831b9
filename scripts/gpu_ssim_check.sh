#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_losses.py tests/test_reference_python.py -m gpu -q -k "ssim" 2>&1 | tail -3
python - <<'PY' 2>&1 | tail -5
import torch, numpy as np, sys
sys.path.insert(0, '.')
from fused_ssim import fused_ssim
from oracle import ref_ssim
a = torch.rand(5, 1, 1500, 1500, device='cuda').requires_grad_(True); b = torch.rand(5, 1, 1500, 1500, device='cuda')
def ev(fn, reps=20):
    fn(); torch.cuda.synchronize(); ts=[]
    for _ in range(reps):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))
def train(f):
    def g():
        a.grad=None; f(a,b).backward()
    return g
fs = ref_ssim.load()
print('ours  train %.3f ms  inference %.3f ms' % (ev(train(fused_ssim)), ev(lambda: fused_ssim(a.detach(), b, train=False))))
print('ref   train %.3f ms  inference %.3f ms' % (ev(train(fs.fused_ssim)), ev(lambda: fs.fused_ssim(a.detach(), b, train=False))))
PY
ncu --set full --clock-control none --import-source on -k regex:ssim_ -c 6 -o gpurun_out/r02_ssim_b python scripts/profile_target.py ssim > gpurun_out/r02_ncu_ssim.log 2>&1; tail -1 gpurun_out/r02_ncu_ssim.log
