#!/bin/bash
set -u
mkdir -p gpurun_out
true

for L in 1 2 1 2; do
VK_COUNT_LANES9=$L timeout 600 python bench.py --workload c3 --steps 50 --warmup 3 --no-cpu-baseline --no-side-legs 2>gpurun_out/r04ae_c3.err | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][-1]
print('c3 lanes9=$L', round(d['value'],1), round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['kernel_ms_per_step'].items()}, round(d['roofline']['frac'],4))"
done
