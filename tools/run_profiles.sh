#!/bin/bash
# Round-end evidence run on one B200 (gpurun): GPU tests + smoke, the bench line of every workload, the reference arm, and --
# each only after its plain command exited 0 -- the ncu launch list and ONE `ncu --set full` capture of a single c2 / c3 step,
# condensed on the box (the reports themselves exceed what gpurun carries back).
set -x
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r2_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r2_smoke.log 2>&1
for w in c2_1280x720_orb5000 c1_640x480_orb5000 c3_1280x720_surf128 c4_window10_orb5000 c5_1920x1200_orb10000; do
  timeout 600 python bench.py --workload $w > $O/r2_bench_$w.json 2> $O/r2_bench_$w.err
done
timeout 600 python bench.py --impl reference > $O/r2_bench_reference_arm.json 2> $O/r2_bench_reference_arm.err
# c2: 18 launches per step, 3 warm-up steps precede the timed ones
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu > $O/r2_plain_c2.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu > $O/r2_ncu_c2_list.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -s 54 -c 18 -f -o /tmp/r2_c2_full python bench.py --steps 2 --warmup 3 --no-cpu > $O/r2_ncu_c2_full.log 2>&1 && \
python tools/ncu_summary.py /tmp/r2_c2_full.ncu-rep > $O/r2_c2_ncu_full_summary.csv && \
python tools/ncu_lines.py /tmp/r2_c2_full.ncu-rep "" 12 > $O/r2_c2_hot_lines.txt 2>&1
# c3: 14 launches per step
timeout 600 python bench.py --workload c3_1280x720_surf128 --steps 2 --warmup 3 --no-cpu > $O/r2_plain_c3.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_c3.csv python bench.py --workload c3_1280x720_surf128 --steps 2 --warmup 3 --no-cpu > $O/r2_ncu_c3_list.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -s 42 -c 14 -f -o /tmp/r2_c3_full python bench.py --workload c3_1280x720_surf128 --steps 2 --warmup 3 --no-cpu > $O/r2_ncu_c3_full.log 2>&1 && \
python tools/ncu_summary.py /tmp/r2_c3_full.ncu-rep > $O/r2_c3_ncu_full_summary.csv && \
python tools/ncu_lines.py /tmp/r2_c3_full.ncu-rep "" 12 > $O/r2_c3_hot_lines.txt 2>&1
ls -la /tmp/*.ncu-rep
tail -3 $O/r2_pytest_gpu.log; tail -1 $O/r2_smoke.log; du -sh $O
