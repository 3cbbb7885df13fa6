#!/bin/bash
# ncu evidence for one workload: launch list (gpu__time_duration) + one --set full capture of the voxelize kernel.
# usage: bash scripts/gpu_profile.sh <workload> <batch> <tag>
WL=${1:-cfg4}; B=${2:-256}; TAG=${3:-r1}
mkdir -p gpurun_out
CMD="python bench.py --workload $WL --steps 2 --warmup 3 --batch $B --no-cpu-baseline"
$CMD > gpurun_out/plain_${WL}_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${WL}_${TAG}.csv $CMD > gpurun_out/ncu_launches_${WL}_${TAG}.log 2>&1
echo "launch-list rc=$?"
$CMD > gpurun_out/plain2_${WL}_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mvx_voxelize -s 3 -c 2 -f -o gpurun_out/prof_${WL}_${TAG} $CMD > gpurun_out/ncu_full_${WL}_${TAG}.log 2>&1
echo "full rc=$?"
ls -la gpurun_out | tail -12
