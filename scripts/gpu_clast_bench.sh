#!/bin/bash
# channels-last bench lines: bash scripts/gpu_clast_bench.sh ["cfg4 cfg2 cfg5"]
for wl in ${1:-cfg4 cfg2 cfg5}; do
  python bench.py --workload $wl --steps 5 --min-seconds 0.3 --molecules 200000 --no-cpu-baseline --channels-last > /tmp/b.json 2>/tmp/b.err || tail -3 /tmp/b.err
  python - "$wl" <<'PY'
import json, sys
d = json.load(open("/tmp/b.json")); r = d["roofline"]
print(f"{sys.argv[1]} channels-last: kernel_ms {r['kernel_ms']:.4f} frac {r['frac']:.3f} value {d['value']:.0f} parity {d['parity'].get('ok')}")
PY
done
