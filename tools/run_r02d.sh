#!/bin/bash
# per-kernel device times (ncu launch list) of the current tree and of the round-1 tree on the same box
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --in-flight 1"
VK_GRAPH=0 $CMD > gpurun_out/r02d_plain.log 2>&1 && VK_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/r02d_launches_now.csv $CMD > gpurun_out/r02d_ncu_now.log 2>&1
echo "now rc=$?"
cd _r01
$CMD > ../gpurun_out/r02d_plain_r01.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file ../gpurun_out/r02d_launches_r01.csv $CMD > ../gpurun_out/r02d_ncu_r01.log 2>&1
echo "r01 rc=$?"
