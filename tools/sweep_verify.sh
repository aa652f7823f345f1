#!/bin/bash
# tuning sweep: launch shape of the pruned cross-check's region-0 verify kernel (FE_VERIFY_VARIANT)
for v in ${VARIANTS:-0 1 2 3 4}; do
  FE_VERIFY_VARIANT=$v python bench.py --steps 8 --warmup 3 --no-cpu --e2e-workers 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
s=[x for x in d['stages'] if x['kernel']=='hamming_cross'][0]
print('variant $v: cross %.3f ms  step %.3f ms  value %.0f  matches %.1f' % (s['ms_per_step'], d['ms_per_step'], d['value'], d['counts']['crosscheck_matches_per_pair_mean']))
"
done
