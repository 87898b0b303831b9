#!/bin/bash
# Round-2 GPU pass 1: GPU tests, launch-shape checksum, both bench arms (reference first, like the driver).
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.log
tail -5 gpurun_out/r02_pytest_gpu.log
python scripts/gpu_checksum.py > gpurun_out/r02_checksum.log 2>&1; tail -20 gpurun_out/r02_checksum.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo "ref rc=$?"
python bench.py > gpurun_out/r02_bench_ours.json 2> gpurun_out/r02_bench_ours.err; echo "ours rc=$?"
tail -c 1500 gpurun_out/r02_bench_ours.err
head -c 3000 gpurun_out/r02_bench_ours.json
