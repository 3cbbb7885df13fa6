#!/bin/bash
# channels-last output: parity tests + bench lines next to the reference layout
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "channels_last or kernel_variants or reduced_precision or no_out_of_bounds" > gpurun_out/pytest_clast.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_clast.log
tail -5 gpurun_out/pytest_clast.log
for wl in cfg4 cfg2 cfg5; do
  for fl in "" "--channels-last"; do
    python bench.py --workload $wl --steps 5 --min-seconds 0.3 --molecules 200000 --no-cpu-baseline $fl > /tmp/b.json 2>/tmp/b.err || tail -3 /tmp/b.err
    cp /tmp/b.json "gpurun_out/bench_${wl}_clast${fl:+1}.json"
    python - "$wl" "$fl" <<'PY'
import json, sys
d = json.load(open("/tmp/b.json")); r = d["roofline"]
print(f"{sys.argv[1]} [{sys.argv[2]}]: kernel_ms {r['kernel_ms']:.4f} frac {r['frac']:.3f} value {d['value']:.0f} e2e {d['e2e']['value']:.0f} parity {d['parity'].get('ok')} err {d['parity'].get('max_abs_err_over_peak')}")
PY
  done
done
