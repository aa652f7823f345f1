#!/bin/bash
# tuning sweep: hamming_cross_kernel launch variants (FE_CROSS_VARIANT), stage time from bench.py
for v in ${VARIANTS:-0 1 2 3 4 5 6}; do
  FE_CROSS_VARIANT=$v python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
s=[x for x in d['stages'] if x['kernel']=='hamming_cross'][0]
print('variant $v: cross %.3f ms  step %.3f ms  value %.0f' % (s['ms_per_step'], d['ms_per_step'], d['value']))
"
done
