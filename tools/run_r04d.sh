#!/bin/bash
set -u
mkdir -p gpurun_out
VK_COUNT_LANES=${LANES:-2} VK_GRAPH=0 timeout 600 ncu --set full --import-source on --clock-control none -k regex:countt -s 2 -c 1 -o gpurun_out/r04j_countt -f python tools/trace_step.py > gpurun_out/r04j_ncu.log 2>&1
echo "ncu rc=$?"
