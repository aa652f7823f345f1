set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu_a.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke_a.log 2>&1
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2_plain_c2.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2_ncu_c2_list.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -s 51 -c 40 -f -o gpurun_out/r2_c2_full_b python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2_ncu_c2_full.log 2>&1
timeout 600 python bench.py --workload c3_1280x720_surf128 --steps 2 --warmup 3 --no-cpu > gpurun_out/r2_plain_c3.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_c3.csv python bench.py --workload c3_1280x720_surf128 --steps 2 --warmup 3 --no-cpu > gpurun_out/r2_ncu_c3_list.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -s 42 -c 34 -f -o gpurun_out/r2_c3_full_b python bench.py --workload c3_1280x720_surf128 --steps 2 --warmup 3 --no-cpu > gpurun_out/r2_ncu_c3_full.log 2>&1
tail -3 gpurun_out/r2_pytest_gpu_a.log; cat gpurun_out/r2_smoke_a.log | tail -2
