#!/bin/bash
set -u
mkdir -p gpurun_out
for L in -1 0; do
VK_COUNT_LANES=$L timeout 900 python bench.py --workload c5 --total-bases 15000000000 --steps 3 2>gpurun_out/r04r_c5_$L.err | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][-1]
print('c5 15Gbp lanes=$L', round(d['value'],1), round(d['ms_per_step'],3), 'fallbacks', d.get('count_fallbacks'), d.get('last_step_timings'))"
done
