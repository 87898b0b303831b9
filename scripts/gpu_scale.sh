#!/bin/bash
# N-GPU run of both bench arms, launched exactly as the driver does (torchrun, one rank per GPU).  usage: gpu_scale.sh N
N=${1:-8}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --impl reference --steps 2 --warmup 1 --configs none > gpurun_out/r02_bench_ref_n$N.json 2> gpurun_out/r02_bench_ref_n$N.err; echo "ref rc=$?"; cut -c1-260 gpurun_out/r02_bench_ref_n$N.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus $N > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "ours rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/r02_bench_n$N.json'))
print('h36m', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d.get('scaling_diag'))
for n,c in d['configs'].items(): print(n, c.get('value'), c.get('e2e',{}).get('value'), c.get('ms_per_step'), c.get('frames_over_capacity_all_ranks'), c.get('total_frames_timed'))
PY
