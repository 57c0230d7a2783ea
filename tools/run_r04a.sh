#!/bin/bash
# countt_kernel (cp.async-staged, one read per lane): parity, then count-kernel time against the default
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "variants" > gpurun_out/r04i_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r04i_pytest.log
for L in 0 2 3 4; do
VK_COUNT_LANES=$L timeout 600 python bench.py --steps 200 --warmup 3 --no-side-legs --no-cpu-baseline --e2e-steps 2 2>gpurun_out/r04i_bench_$L.err | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][-1]
print('lanes=$L', round(d['value'],1), {k:round(v,4) for k,v in d['kernel_ms_per_step'].items()})"
done
