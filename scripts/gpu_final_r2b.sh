#!/bin/bash
# Round-2 final evidence on one B200: bench lines (default incl. cpu_baseline, reference arm, augment, other workloads, output
# variants) and ncu launch lists + full captures of the three workloads.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
python bench.py > gpurun_out/bench_cfg4_default.json 2> gpurun_out/bench_cfg4_default.err; echo "default rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_arm.json 2>/dev/null; echo "reference rc=$?"
python bench.py --augment --no-cpu-baseline > gpurun_out/bench_cfg4_augment.json 2>/dev/null; echo "augment rc=$?"
python bench.py --no-overlap --no-cpu-baseline > gpurun_out/bench_cfg4_no_overlap.json 2>/dev/null; echo "no-overlap rc=$?"
python bench.py --out-dtype bfloat16 --no-cpu-baseline > gpurun_out/bench_cfg4_bf16.json 2>/dev/null; echo "bf16 rc=$?"
python bench.py --channels-last --no-cpu-baseline > gpurun_out/bench_cfg4_channels_last.json 2>/dev/null; echo "clast rc=$?"
for wl in cfg3 cfg2 cfg5; do
  python bench.py --workload $wl --no-cpu-baseline > gpurun_out/bench_$wl.json 2>/dev/null; echo "$wl rc=$?"
done
bash scripts/gpu_profile.sh cfg4 1024 r2b > /dev/null 2>&1
bash scripts/gpu_profile.sh cfg2 256 r2b > /dev/null 2>&1
bash scripts/gpu_profile.sh cfg5 32 r2b > /dev/null 2>&1
ls gpurun_out | head -50
