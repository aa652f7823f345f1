set -x
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
timeout 600 python bench.py --workload c3_1280x720_surf128 > $O/r2_bench_c3_1280x720_surf128.json 2> $O/r2_bench_c3_1280x720_surf128.err
timeout 600 python bench.py --impl reference --workload c3_1280x720_surf128 > $O/r2_bench_reference_arm_c3.json 2> $O/r2_bench_reference_arm_c3.err
timeout 600 python bench.py --workload c3_1280x720_surf128 --steps 2 --warmup 3 --no-cpu > $O/r2_plain_c3.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_c3.csv python bench.py --workload c3_1280x720_surf128 --steps 2 --warmup 3 --no-cpu > $O/r2_ncu_c3_list.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -s 42 -c 14 -f -o /tmp/r2_c3_full python bench.py --workload c3_1280x720_surf128 --steps 2 --warmup 3 --no-cpu > $O/r2_ncu_c3_full.log 2>&1 && \
python tools/ncu_summary.py /tmp/r2_c3_full.ncu-rep > $O/r2_c3_ncu_full_summary.csv && \
python tools/ncu_lines.py /tmp/r2_c3_full.ncu-rep "" 12 > $O/r2_c3_hot_lines.txt 2>&1
timeout 400 python -m pytest tests -m gpu -x -q > $O/r2_pytest_gpu.log 2>&1; tail -2 $O/r2_pytest_gpu.log
