# tuning sweep of the pruned cross-check's class splits / launch shapes (stage times from tools/stage_time.py; the match counts
# at the end of every line must not change)
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
{
for kv in "A=0" "FE_CX_SWAP=0" "FE_CX_T2=48" "FE_CX_T2=36" "FE_CX_T=20" "FE_CX_T=28" "FE_CROSS_MIH=0"; do
  echo "$kv"; env $kv python tools/stage_time.py c2 10
done
} > $O/r2b_sweep_classes3.log 2>&1
cat $O/r2b_sweep_classes3.log
