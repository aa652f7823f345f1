# tuning sweep of the pruned cross-check's class splits / launch shapes (stage times from tools/stage_time.py; the match counts
# at the end of every line must not change) + a per-kernel launch list of the stage
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
{
for kv in "A=0" "FE_CX_ZC=48 FE_CX_ZD=32" "FE_CX_ZC=24 FE_CX_ZD=48"; do
  echo "$kv"; env $kv python tools/stage_time.py c2 10
done
} > $O/r2b_sweep_classes2.log 2>&1
cat $O/r2b_sweep_classes2.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2b_cross_launches.csv -k regex:'verify|mih|classify|finalize|band' -c 80 python tools/stage_time.py c2 1 > $O/r2b_ncu.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2b_cross_launches.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: h=r; s=i; break
k,v,g=h.index('Kernel Name'),h.index('Metric Value'),h.index('Grid Size')
for r in rows[s+1:][-14:]:
    if len(r)>v: print("%-60s %9s ns %s"%(r[k][:60],r[v],r[g]))
PY
