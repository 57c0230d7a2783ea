#!/bin/bash
# Runs ON the GPU box: ncu --set full of ONE launch of the count kernel named by $1 (regex), env passed through.
# usage: VK_COUNT_LANES=1 tools/profile_count.sh countu_kernel r03a
set -u
K=${1:-count_kernel}; TAG=${2:-r03}
mkdir -p gpurun_out
CMD="python tools/trace_step.py"
$CMD > gpurun_out/plain_${TAG}.log 2>&1 || { tail -5 gpurun_out/plain_${TAG}.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"$K" -s 3 -c 1 -o gpurun_out/prof_${TAG} -f $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "full rc=$?"; tail -2 gpurun_out/ncu_full_${TAG}.log
ncu -i gpurun_out/prof_${TAG}.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_${TAG}.ncu-rep --page source --csv > gpurun_out/prof_${TAG}_source.csv 2>/dev/null
ls -la gpurun_out/prof_${TAG}*
