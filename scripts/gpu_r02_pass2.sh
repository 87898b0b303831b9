#!/bin/bash
# Round-2 GPU pass 2: all GPU tests (no -x), then regenerate the optimiser goldens (64 frames/config + the 8-view rig).
mkdir -p gpurun_out/golden
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu2.log
grep -E "passed|failed|FAILED|ERROR" gpurun_out/r02_pytest_gpu2.log | tail -30
python tests/golden/make_golden.py --out gpurun_out/golden > gpurun_out/r02_make_golden.log 2>&1; echo "golden rc=$?"
grep "wrote" gpurun_out/r02_make_golden.log; tail -3 gpurun_out/r02_make_golden.log
