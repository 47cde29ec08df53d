#!/bin/bash
# how does the specialised kernel's throughput scale with resident warps per SM?
for wl in c3 c2; do
  cells=1048576; [ $wl = c2 ] && cells=8388608
  for mb in 1 2 3 4 8; do
    echo -n "$wl maxblocks=$mb: "
    PFRX_SPEC_MAXBLOCKS=$mb python bench.py --workload $wl --cells $cells --steps 2 --warmup 1 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); k=d['config']['kernel']
print(k['threads'],'thr x',k['blocks_per_sm'],'blk/SM', round(d['ms_per_step'],2),'ms', '%.3g'%d['value'],'cells/s')"
  done
done
