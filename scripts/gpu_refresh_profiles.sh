#!/bin/bash
# The round-end GPU pass (round 2), in stages so that one gpurun call stays inside its 64 MiB output limit:
#   STAGE=tests    GPU tests, smoke(), golden regeneration (GOLDEN=1), both bench arms (reference first, like the driver), the
#                  launch list of the bench command
#   STAGE=ncu      the `ncu --set full` captures, summarised ON THE BOX into gpurun_out/profiles_r02/ (scripts/summarize_ncu.py,
#                  scripts/make_traffic_json.py); the .ncu-rep files (~90 MB) are deleted afterwards
#   STAGE=sanitize compute-sanitizer memcheck / racecheck / synccheck over every kernel
R=${ROUND:-r02}
STAGE=${STAGE:-tests}
mkdir -p gpurun_out/golden gpurun_out/profiles_${R}
if [ "$STAGE" = "tests" ]; then
  timeout 900 python -m pytest tests -q -m gpu > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${R}_pytest_gpu.log; tail -3 gpurun_out/${R}_pytest_gpu.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/${R}_smoke.log 2>&1; tail -1 gpurun_out/${R}_smoke.log
  if [ "${GOLDEN:-0}" = "1" ]; then
    timeout 1500 python tests/golden/make_golden.py --out gpurun_out/golden > gpurun_out/${R}_make_golden.log 2>&1; grep "wrote opt" gpurun_out/${R}_make_golden.log
  fi
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_ref.json 2> gpurun_out/${R}_bench_ref.err; cut -c1-200 gpurun_out/${R}_bench_ref.json
  timeout 900 python bench.py > gpurun_out/${R}_bench_ours.json 2> gpurun_out/${R}_bench_ours.err; cut -c1-300 gpurun_out/${R}_bench_ours.json
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${R}_launches_bench.csv python bench.py --steps 2 --warmup 1 --configs none > gpurun_out/${R}_bench_under_ncu.log 2>&1
fi
if [ "$STAGE" = "ncu" ]; then
  P=gpurun_out/profiles_${R}
  for cfg in h36m h36m-occ panoptic occlusion-person-8v; do
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:optimize_kernel -c 1 -o gpurun_out/${R}_opt_${cfg} -f python scripts/profile_target.py opt 2048 ${cfg} > gpurun_out/${R}_ncu_opt_${cfg}.log 2>&1
    python scripts/summarize_ncu.py gpurun_out/${R}_opt_${cfg}.ncu-rep $P/${R}_optimize_kernel_${cfg}_ncu.txt > /dev/null 2>&1
  done
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:fill_zero|render_active|render_bwd_kernel|bin_kernel|gauss_bwd_kernel" -c 5 -o gpurun_out/${R}_dense -f python scripts/profile_target.py dense > gpurun_out/${R}_ncu_dense.log 2>&1
  python scripts/summarize_ncu.py gpurun_out/${R}_dense.ncu-rep $P/${R}_dense_rasterizer_ncu.txt > /dev/null 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:ssim_" -c 6 -o gpurun_out/${R}_ssim -f python scripts/profile_target.py ssim > gpurun_out/${R}_ncu_ssim.log 2>&1
  python scripts/summarize_ncu.py gpurun_out/${R}_ssim.ncu-rep $P/${R}_ssim_ncu.txt > /dev/null 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:loss_" -c 4 -o gpurun_out/${R}_loss -f python scripts/profile_target.py loss > gpurun_out/${R}_ncu_loss.log 2>&1
  python scripts/summarize_ncu.py gpurun_out/${R}_loss.ncu-rep $P/${R}_loss_ncu.txt > /dev/null 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:roi_|dlt_" -c 4 -o gpurun_out/${R}_setup -f python scripts/profile_target.py setup > gpurun_out/${R}_ncu_setup.log 2>&1
  python scripts/summarize_ncu.py gpurun_out/${R}_setup.ncu-rep $P/${R}_setup_ncu.txt > /dev/null 2>&1
  python scripts/make_traffic_json.py ${R} $P > /dev/null 2>&1
  ls -la $P
  rm -f gpurun_out/${R}_*.ncu-rep
fi
if [ "$STAGE" = "ncu_one" ]; then        # re-capture of the fused optimiser for ONE config (CFG=...); merge its entry into profiles/*_traffic.json by hand
  P=gpurun_out/profiles_${R}
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:optimize_kernel -c 1 -o gpurun_out/${R}_opt_${CFG} -f python scripts/profile_target.py opt 2048 ${CFG} > gpurun_out/${R}_ncu_opt_${CFG}.log 2>&1
  python scripts/summarize_ncu.py gpurun_out/${R}_opt_${CFG}.ncu-rep $P/${R}_optimize_kernel_${CFG}_ncu.txt > /dev/null 2>&1
  python scripts/make_traffic_json.py ${R} $P > /dev/null 2>&1
  rm -f gpurun_out/${R}_*.ncu-rep
fi
if [ "$STAGE" = "ncu_ssim" ]; then
  P=gpurun_out/profiles_${R}
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:ssim_" -c 6 -o gpurun_out/${R}_ssim -f python scripts/profile_target.py ssim > gpurun_out/${R}_ncu_ssim.log 2>&1
  python scripts/summarize_ncu.py gpurun_out/${R}_ssim.ncu-rep $P/${R}_ssim_ncu.txt > /dev/null 2>&1
  rm -f gpurun_out/${R}_*.ncu-rep
fi
if [ "$STAGE" = "ncu_setup" ]; then      # re-capture of the setup kernels only (cheap)
  P=gpurun_out/profiles_${R}
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:roi_|dlt_" -c 4 -o gpurun_out/${R}_setup -f python scripts/profile_target.py setup > gpurun_out/${R}_ncu_setup.log 2>&1
  python scripts/summarize_ncu.py gpurun_out/${R}_setup.ncu-rep $P/${R}_setup_ncu.txt > /dev/null 2>&1
  rm -f gpurun_out/${R}_*.ncu-rep
fi
if [ "$STAGE" = "sanitize" ]; then
  for tool in memcheck racecheck synccheck; do
    timeout 1200 compute-sanitizer --tool $tool python scripts/gpu_sanitize_target.py 4 > gpurun_out/${R}_sanitize_$tool.log 2>&1; tail -2 gpurun_out/${R}_sanitize_$tool.log
  done
fi
du -sh gpurun_out
