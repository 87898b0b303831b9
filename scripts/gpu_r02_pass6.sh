#!/bin/bash
# Round-2 GPU pass 6: bench both arms with the factored-GT kernel + ncu captures of the fused optimiser.
mkdir -p gpurun_out
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref6.json 2> gpurun_out/r02_bench_ref6.err; echo "ref rc=$?"
python bench.py > gpurun_out/r02_bench_ours6.json 2> gpurun_out/r02_bench_ours6.err; echo "ours rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_ours6.json'))
print('h36m', d['value'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], 'acc', d['accuracy']['vs_reference'])
for n,c in d['configs'].items(): print(n, c.get('value'), c.get('e2e',{}).get('value'), c.get('frames_over_capacity_all_ranks'), c.get('accuracy',{}).get('vs_reference'))
print('ssim', d['dense_surface'].get('fused_ssim_5x1x1500x1500'))
print('m2', d['m2_rasterizer_dense']['value'], d['m2_rasterizer_dense']['roofline']['frac'])
PY
ncu --set full --clock-control none --import-source on -k regex:optimize_kernel -c 1 -o gpurun_out/r02_opt_h36m python scripts/profile_target.py opt 2048 h36m > gpurun_out/r02_ncu_opt.log 2>&1; tail -1 gpurun_out/r02_ncu_opt.log
ncu --set full --clock-control none --import-source on -k regex:optimize_kernel -c 1 -o gpurun_out/r02_opt_panoptic python scripts/profile_target.py opt 1024 panoptic > gpurun_out/r02_ncu_opt_p.log 2>&1; tail -1 gpurun_out/r02_ncu_opt_p.log
