#!/bin/bash
# Round-2 GPU pass 5: factored GT profiles in the fused kernel: tests, capacities, speed.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_gpu5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu5.log
grep -E "passed|failed|FAILED|ERROR|rc=" gpurun_out/r02_pytest_gpu5.log | tail -12
python scripts/gpu_capacity_scan.py h36m 320 2048 0,-1 2>&1 | tail -8
python scripts/gpu_capacity_scan.py occlusion-person-8v 512,416,384 2048 0,-1 2>&1 | tail -14
python scripts/gpu_capacity_scan.py panoptic 1024 2048 0,-1 2>&1 | tail -8
python scripts/gpu_capacity_scan.py h36m-occ 320 2048 0 2>&1 | tail -3
