#!/bin/bash
set -u
mkdir -p gpurun_out
for CFG in "2 0x1" "2 0x3" "3 0x1" "4 0x1" "4 0x3" "2 0x71"; do
set -- $CFG
VK_COUNTT_KNOBS=$2 VK_COUNT_LANES=$1 timeout 600 python bench.py --steps 200 --warmup 3 --no-side-legs --no-cpu-baseline --e2e-steps 2 2>gpurun_out/r04l_bench.err | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][-1]
print('lanes=$1 knobs=$2', round(d['value'],1), {k:round(v,4) for k,v in d['kernel_ms_per_step'].items()})"
done
VK_COUNTT_KNOBS=0x1 VK_COUNT_LANES=2 VK_GRAPH=0 timeout 600 ncu --set full --import-source on --clock-control none -k regex:countt -s 2 -c 1 -o gpurun_out/r04l_countt -f python tools/trace_step.py > gpurun_out/r04l_ncu.log 2>&1
echo "ncu rc=$?"
