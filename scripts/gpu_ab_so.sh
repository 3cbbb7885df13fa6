#!/bin/bash
# A/B of library builds on the dense workloads: usage: bash scripts/gpu_ab_so.sh "so1 so2 ..." ["cfg2 cfg5"] [extra bench flags]
SOS=$1; WLS=${2:-"cfg2 cfg5"}; shift 2
for wl in $WLS; do
  for so in $SOS; do
    if [ "$so" = "default" ]; then unset MVX_SO; else export MVX_SO=$PWD/$so; fi
    python bench.py --workload $wl --steps 5 --min-seconds 0.3 --no-cpu-baseline "$@" > /tmp/ab.json 2>/tmp/ab.err || tail -3 /tmp/ab.err
    python - "$wl" "$so" <<'PY'
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
