#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "countt9 or variants or config3 or fire_and_forget" > gpurun_out/r04af_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r04af_pytest.log | cut -c1-200
for L in -1 0; do
VK_COUNT_LANES9=$L timeout 600 python bench.py --workload c3 --steps 50 --warmup 3 --no-cpu-baseline --no-side-legs 2>gpurun_out/r04af_c3.err | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][-1]
print('c3 lanes9=$L', round(d['value'],1), round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['kernel_ms_per_step'].items()}, round(d['roofline']['frac'],4), d['roofline']['kernel'])"
done
