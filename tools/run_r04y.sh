#!/bin/bash
set -u
mkdir -p gpurun_out
for T in 3 4 5 6 8; do
timeout 600 python bench.py --steps 1000 --warmup 3 --no-side-legs --no-cpu-baseline --e2e-steps 2 --in-flight $T 2>gpurun_out/r04y_bench.err | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][-1]
print('in-flight=$T', round(d['value'],1), round(d['ms_per_step'],4), round(d['ms_per_step_wall'],4))"
done
