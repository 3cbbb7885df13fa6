#!/bin/bash
# Kernel-form comparison on the dense workloads: MVX_KERNEL override x workload.  Output: gpurun_out/forms.txt
mkdir -p gpurun_out
: > gpurun_out/forms.txt
for wl in ${WORKLOADS:-cfg2 cfg5}; do
  for var in ${VARIANTS:-cells tiles pipe}; do
    MVX_KERNEL=$var timeout 300 python bench.py --workload $wl --steps ${STEPS:-20} --warmup 3 --no-cpu-baseline 2> gpurun_out/err_${wl}_${var}.log | python -c "
import json,sys
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); r=d['roofline']
        print('$wl $var value=%.0f e2e=%.0f GBps=%.0f frac=%.3f vox_ms=%.3f bin_ms=%.3f prep_ms=%.3f clk=%s' % (d['value'], d['e2e']['value'], r['achieved'], r['frac'], r['kernel_ms'], r['step_share']['bin_ms'], r['step_share']['prep_ms'], d['clocks']['sm_mhz']))
" >> gpurun_out/forms.txt
  done
done
cat gpurun_out/forms.txt
