#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "variants or epochs" > gpurun_out/r04u_pytest.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r04u_pytest.log
VK_N=15000000000 python tools/diag_big.py 2>&1 | tail -3
VK_COUNT_LANES=-1 timeout 600 python bench.py --steps 200 --warmup 3 --no-side-legs --no-cpu-baseline --e2e-steps 2 2>gpurun_out/r04u_bench.err | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][-1]
print('c2', round(d['value'],1), round(d['roofline']['frac'],4), d['roofline']['kernel'], {k:round(v,4) for k,v in d['kernel_ms_per_step'].items()})"
