#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "epochs" > gpurun_out/r04v_pytest.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r04v_pytest.log
for KN in 0 0x2 0 0x2 0x4000; do
VK_COUNTT_KNOBS=$KN VK_COUNT_LANES=2 timeout 600 python bench.py --steps 200 --warmup 3 --no-side-legs --no-cpu-baseline --e2e-steps 2 2>gpurun_out/r04v_bench.err | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][-1]
print('knobs=$KN', round(d['value'],1), round(d['roofline']['frac'],4), {k:round(v,4) for k,v in d['kernel_ms_per_step'].items()})"
done
