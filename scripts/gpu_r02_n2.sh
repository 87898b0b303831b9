#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --configs none > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_n2.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('scaling_diag'))
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 --configs none > gpurun_out/r02_bench_ref_n2.json 2> gpurun_out/r02_bench_ref_n2.err; echo "rc=$?"; head -c 600 gpurun_out/r02_bench_ref_n2.json
