#!/bin/bash
# Tile z extent sweep (MVX_TZ) on the dense workloads.  Output: gpurun_out/tz.txt
mkdir -p gpurun_out
: > gpurun_out/tz.txt
for wl in cfg2 cfg5; do
  for tz in ${TZS:-default 16 32 48 64}; do
    if [ $tz = default ]; then unset MVX_TZ; else export MVX_TZ=$tz; fi
    timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/err_tz.log | python -c "
import json,sys
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); r=d['roofline']
        print('$wl tz=$tz value=%.0f vox_ms=%.3f bin_ms=%.3f prep_ms=%.3f' % (d['value'], r['kernel_ms'], r['step_share']['bin_ms'], r['step_share']['prep_ms']))
" >> gpurun_out/tz.txt
  done
done
cat gpurun_out/tz.txt
