# tuning sweep of the pruned cross-check (stage times from tools/stage_time.py; the match counts at the end of every line must not change)
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
{
for kv in "FE_CX_OVERLAP=0" "A=0" "FE_CX_JOIN1=1"; do
  echo "$kv"; env $kv python tools/stage_time.py c2 10
done
} > $O/r2b_sweep_classes4.log 2>&1
cat $O/r2b_sweep_classes4.log
