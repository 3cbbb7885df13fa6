#!/bin/bash
# usage: ENVS="MVX_CH=16 MVX_CH=8" WORKLOADS="cfg2 cfg5" bash scripts/gpu_envsweep.sh   (each ENVS token is one VAR=VAL[,VAR=VAL] set)
mkdir -p gpurun_out; : > gpurun_out/envsweep.txt
if [ -n "$RUN_TESTS" ]; then python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -3; fi
for wl in ${WORKLOADS:-cfg2}; do
  for e in ${ENVS:-none}; do
    ( IFS=','; for kv in $e; do [ "$kv" != none ] && export "$kv"; done
      python bench.py --workload $wl --steps ${STEPS:-20} --warmup 3 --no-cpu-baseline 2> gpurun_out/err_env.log | python -c "
import json,sys
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); r=d['roofline']
        print('$wl $e value=%.0f e2e=%.0f GBps=%.0f frac=%.3f vox_ms=%.3f bin_ms=%.3f' % (d['value'], d['e2e']['value'], r['achieved'], r['frac'], r['kernel_ms'], r['step_share']['bin_ms']))
" >> gpurun_out/envsweep.txt; tail -2 gpurun_out/err_env.log | grep -i error >> gpurun_out/envsweep.txt )
  done
done
cat gpurun_out/envsweep.txt
