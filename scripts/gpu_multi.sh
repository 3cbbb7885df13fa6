#!/bin/bash
# Multi-GPU evidence on one box: usage: bash scripts/gpu_multi.sh N   (run under gpurun --gpus N)
# cfg4 sweep (sharded by molecule index) and cfg2 (host-fed dense batches: e2e scaling) at N GPUs, with the optional NCCL
# all-gather of finished grids timed separately.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 > gpurun_out/gpus_n$N.txt
for wl in cfg4 cfg2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700+N)) bench.py --gpus $N --workload $wl --gather --no-cpu-baseline > gpurun_out/scale_${wl}_n$N.json 2> gpurun_out/scale_${wl}_n$N.err
  echo "$wl n=$N rc=$?"
  python - "$wl" "$N" <<'PY'
import json, sys
wl, n = sys.argv[1], sys.argv[2]
for ln in open(f"gpurun_out/scale_{wl}_n{n}.json"):
    if ln.startswith("{"):
        d = json.loads(ln)
        print(wl, "n_gpus", d["n_gpus"], "value %.0f" % d["value"], "e2e %.0f" % d["e2e"]["value"], "frac %.3f" % d["roofline"]["frac"],
              "gather", d["e2e"].get("gather") or d.get("gather"), d["clocks"])
PY
done
