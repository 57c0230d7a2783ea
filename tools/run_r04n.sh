#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "variants" > gpurun_out/r04p_pytest.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r04p_pytest.log
for CFG in "2 0" "-1 0"; do
set -- $CFG
VK_COUNTT_KNOBS=$2 VK_COUNT_LANES=$1 timeout 600 python bench.py --steps 200 --warmup 3 --no-side-legs --no-cpu-baseline --e2e-steps 2 2>gpurun_out/r04p_bench.err | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][-1]
print('lanes=$1 knobs=$2', round(d['value'],1), round(d['roofline']['frac'],4), d['roofline']['kernel'], {k:round(v,4) for k,v in d['kernel_ms_per_step'].items()})"
done
for L in -1 0; do
VK_COUNT_LANES=$L python bench.py --workload c4 --in-flight 8 --steps 3 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][-1]
print('c4 lanes=$L', round(d['value'],1), round(d['ms_per_step'],3))"
done
