#!/bin/bash
# Kernel-variant sweep: MVX_KERNEL / MVX_LPR overrides across workloads.  Output: gpurun_out/sweep.txt
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
: > gpurun_out/sweep.txt
for wl in ${WORKLOADS:-cfg4 cfg3 cfg2 cfg5}; do
  for var in ${VARIANTS:-rows cells2 cells4 cells16}; do
    case $var in
      rows) export MVX_KERNEL=rows; unset MVX_LPR;;
      cells*) export MVX_KERNEL=cells; export MVX_LPR=${var#cells};;
    esac
    python bench.py --workload $wl --steps ${STEPS:-20} --warmup 3 --no-cpu-baseline 2> gpurun_out/err_${wl}_${var}.log | python -c "
import json,sys
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); r=d['roofline']
        print('$wl $var value=%.0f e2e=%.0f GBps=%.0f frac=%.3f vox_ms=%.3f bin_ms=%.3f prep_ms=%.3f clk=%s' % (d['value'], d['e2e']['value'], r['achieved'], r['frac'], r['kernel_ms'], r['step_share']['bin_ms'], r['step_share']['prep_ms'], d['clocks']['sm_mhz']))
" >> gpurun_out/sweep.txt
  done
done
cat gpurun_out/sweep.txt
