#!/bin/bash
# Round-2 GPU pass 3: tests, optimiser experiments (upper bounds: no GT loads / fast exp), outlier diagnosis, both bench arms.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu3.log
grep -E "passed|failed|FAILED|ERROR" gpurun_out/r02_pytest_gpu3.log | tail -30
python scripts/gpu_tune_opt.py --variants ";SSB_EXP_NOGT=1;SSB_EXP_FASTEXP=1;SSB_EXP_NOGT=1,SSB_EXP_FASTEXP=1" > gpurun_out/r02_tune_exp.log 2>&1; cat gpurun_out/r02_tune_exp.log | cut -c1-220
python scripts/gpu_outlier_diag.py occlusion-person-8v > gpurun_out/r02_outlier.log 2>&1; tail -5 gpurun_out/r02_outlier.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref3.json 2> gpurun_out/r02_bench_ref3.err; echo "ref rc=$?"
python bench.py > gpurun_out/r02_bench_ours3.json 2> gpurun_out/r02_bench_ours3.err; echo "ours rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_ours3.json')); r=json.load(open('gpurun_out/r02_bench_ref3.json'))
print('value',d['value'],'e2e',d['e2e']['value'])
print('ssim ours',d['dense_surface'].get('fused_ssim_5x1x1500x1500'))
print('ssim ref',r.get('fused_ssim_5x1x1500x1500'))
print('dropin',d['dense_surface'].get('dropin_loop_frames_per_s'))
PY
