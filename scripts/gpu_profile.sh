#!/bin/bash
# ncu evidence for one workload: launch list (gpu__time_duration) + one --set full capture of the voxelize kernel,
# summarised ON THE BOX (raw page, per-block instruction/stall histogram, hottest SASS lines); the .ncu-rep itself is
# dropped so that gpurun_out/ stays under its size limit.
# usage: bash scripts/gpu_profile.sh <workload> <batch> <tag> [kernel regex]
WL=${1:-cfg4}; B=${2:-256}; TAG=${3:-r1}; KRE=${4:-mvx_voxelize_(pipe|cells|tiles)}
mkdir -p gpurun_out
# a short run: 6 chunks of the sweep / a 5 ms pool loop, no parity / CPU legs (ncu replays every kernel ~40 times)
CMD="python bench.py --workload $WL --steps 2 --warmup 3 --batch $B --molecules $((B * 6)) --min-seconds 0.005 --no-cpu-baseline --no-parity"
$CMD > gpurun_out/plain_${WL}_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/launches_${WL}_${TAG}.csv $CMD > /dev/null 2>&1
echo "launch-list rc=$?"
$CMD > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"$KRE" -s 3 -c 1 -f -o /tmp/prof_${WL}_${TAG} $CMD > gpurun_out/ncu_full_${WL}_${TAG}.log 2>&1
echo "full rc=$?"
python scripts/ncu_summary.py /tmp/prof_${WL}_${TAG}.ncu-rep 60 > gpurun_out/ncu_summary_${WL}_${TAG}.txt 2>&1
python scripts/ncu_hot.py /tmp/prof_${WL}_${TAG}.ncu-rep 40 > gpurun_out/ncu_hot_${WL}_${TAG}.txt 2>&1
ncu -i /tmp/prof_${WL}_${TAG}.ncu-rep --page raw --csv > gpurun_out/ncu_full_${WL}_${TAG}_raw.csv 2>/dev/null
rm -f /tmp/prof_${WL}_${TAG}.ncu-rep
ls -la gpurun_out | tail -8
