#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "variants or full_size" > gpurun_out/r04s_pytest.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r04s_pytest.log
VK_COUNT_LANES=-1 timeout 600 python bench.py --steps 200 --warmup 3 --no-side-legs --no-cpu-baseline --e2e-steps 2 2>gpurun_out/r04s_bench.err | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][-1]
print('c2', round(d['value'],1), round(d['roofline']['frac'],4), d['roofline']['kernel'], {k:round(v,4) for k,v in d['kernel_ms_per_step'].items()})"
for L in -1 0; do
VK_COUNT_LANES=$L timeout 900 python bench.py --workload c5 --total-bases 15000000000 --steps 3 2>gpurun_out/r04s_c5_$L.err | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][-1]
print('c5 15Gbp lanes=$L', round(d['value'],1), round(d['ms_per_step'],3))"
done
