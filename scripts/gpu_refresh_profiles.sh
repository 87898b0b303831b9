set -x
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu3.log 2>&1; tail -2 gpurun_out/pytest_gpu3.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/smoke3.log 2>&1; tail -1 gpurun_out/smoke3.log
timeout 600 python bench.py > gpurun_out/bench_ours3.json 2> gpurun_out/bench_ours3.err; cut -c1-400 gpurun_out/bench_ours3.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref3.json 2> gpurun_out/bench_ref3.err; cut -c1-300 gpurun_out/bench_ref3.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01c.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:optimize_kernel -c 1 -o gpurun_out/prof_opt_r01e -f python scripts/profile_target.py opt 2048 > gpurun_out/ncu_opt_e.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:fill_zero|render_active|render_bwd_kernel|bin_kernel|gauss_bwd_kernel" -c 5 -o gpurun_out/prof_dense_r01e -f python scripts/profile_target.py dense > gpurun_out/ncu_dense_e.log 2>&1
ls -la gpurun_out/*r01e* gpurun_out/launches_r01c.csv
