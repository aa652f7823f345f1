#!/bin/bash
# tuning sweep: chunk size of the overlapped pipeline x number of host worker contexts -> e2e pairs/s
for c in ${CHUNKS:-16 24 32 48 96}; do for wk in ${WORKERS:-1 2 3}; do
  FE_CHUNK_PAIRS=$c python bench.py --steps 12 --warmup 3 --no-cpu --e2e-workers $wk 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('chunk $c workers $wk: e2e %.0f pairs/s  value %.0f' % (d['e2e']['value'], d['value']))
"
done; done
