#!/bin/bash
set -u
mkdir -p gpurun_out
for X in 0 -4 -8 -16 4; do
VK_COUNT_EXTRA=$X timeout 600 python bench.py --steps 1000 --warmup 3 --no-side-legs --no-cpu-baseline --e2e-steps 2 2>gpurun_out/r04ab_bench.err | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][-1]
print('extra=$X', round(d['value'],1), round(d['ms_per_step'],4), round(d['kernel_ms_per_step']['count'],4))"
done
