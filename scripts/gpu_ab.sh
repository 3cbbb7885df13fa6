#!/bin/bash
# A/B of a kernel switch on the dense workloads: kernel_ms, step value and parity of bench.py for each setting of an
# environment variable.  usage: bash scripts/gpu_ab.sh VAR "v1 v2 ..." ["cfg2 cfg5"]
VAR=$1; VALS=$2; WLS=${3:-"cfg2 cfg5"}
for wl in $WLS; do
  for v in $VALS; do
    env $VAR=$v python bench.py --workload $wl --steps 5 --min-seconds 0.3 --no-cpu-baseline > /tmp/ab.json 2>/tmp/ab.err || tail -3 /tmp/ab.err
    python - "$wl" "$VAR=$v" <<'PY'
import json, sys
try:
    d = json.load(open("/tmp/ab.json"))
    r = d["roofline"]
    print(f"{sys.argv[1]} {sys.argv[2]}: kernel_ms {r['kernel_ms']:.4f} frac {r['frac']:.3f} bin_ms {r['step_share']['bin_ms']:.4f} prep_ms {r['step_share']['prep_ms']:.4f} value {d['value']:.0f} e2e {d['e2e']['value']:.0f} parity {(d.get('parity') or {}).get('ok')} err {(d.get('parity') or {}).get('max_abs_err_over_peak')}")
except Exception as e:
    print(sys.argv[1], sys.argv[2], "FAILED", e)
PY
  done
done
