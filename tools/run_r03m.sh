#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "more_levels or config4 or variants or stage_functions or query or golden" > gpurun_out/r03m_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r03m_pytest.log
for T in 8; do
python bench.py --workload c4 --in-flight $T --steps 3 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][-1]
print('c4 T=$T', round(d['value'],1), round(d['ms_per_step'],3))"
done
python bench.py --steps 400 --warmup 3 --no-side-legs --no-cpu-baseline --e2e-steps 2 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][-1]
print('c2', round(d['value'],1), {k:round(v,4) for k,v in d['kernel_ms_per_step'].items()})"
