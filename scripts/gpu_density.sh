#!/bin/bash
# Kernel-form selection threshold: cfg2 geometry at several atom counts, cells vs pipe.  Output: gpurun_out/density.txt
mkdir -p gpurun_out
: > gpurun_out/density.txt
for atoms in ${ATOMS:-125 250 500 1000 2000}; do
  for var in cells pipe; do
    MVX_KERNEL=$var timeout 300 python bench.py --workload cfg2 --atoms $atoms --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/err_density.log | python -c "
import json,sys
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); r=d['roofline']
        print('atoms=$atoms $var value=%.0f vox_ms=%.3f bin_ms=%.3f prep_ms=%.3f' % (d['value'], r['kernel_ms'], r['step_share']['bin_ms'], r['step_share']['prep_ms']))
" >> gpurun_out/density.txt
  done
done
cat gpurun_out/density.txt
