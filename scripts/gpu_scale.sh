#!/bin/bash
# Scaling run as the driver does it: N = 1, 2, 4, 8 back to back.  Output: gpurun_out/scale_*.json
mkdir -p gpurun_out
for n in ${NS:-1 2 4 8}; do
  if [ "$n" = 1 ]; then
    python bench.py --gpus 1 --steps ${STEPS:-200} --warmup 3 --no-cpu-baseline > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps ${STEPS:-200} --warmup 3 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  fi
  echo "n=$n rc=$?"
  python -c "
import json
for ln in open('gpurun_out/scale_n$n.json'):
    if ln.startswith('{'):
        d=json.loads(ln); print('n_gpus',d['n_gpus'],'value %.0f'%d['value'],'e2e %.0f'%d['e2e']['value'],'ms/step %.3f'%d['ms_per_step'],'frac %.3f'%d['roofline']['frac'], d['clocks'])
"
done
