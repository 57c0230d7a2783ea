#!/bin/bash
# GPU run: full GPU suite, then A/B of the step variants (graph / plain launches, packed / text count kernels)
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
tail -15 gpurun_out/r02a_pytest.log
for PK in 1 0; do for GR in 1 0; do
  VK_PACKED=$PK VK_GRAPH=$GR timeout 300 python bench.py --steps 400 --warmup 3 --no-cpu-baseline --e2e-steps 4 > gpurun_out/r02a_bench_pk${PK}_gr${GR}.json 2> gpurun_out/r02a_bench_pk${PK}_gr${GR}.err
  echo "pk=$PK gr=$GR rc=$?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02a_bench_pk${PK}_gr${GR}.json"))
    print({k:d[k] for k in ("value","ms_per_step","ms_per_step_wall","gpu_launches")}, d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["kernel_ms_per_step"], d["one_context"]["ms_per_step_device"], d["one_context"]["ms_per_step_wall"], d["e2e"]["value"])
except Exception as e:
    print("ERR", e); print(open("gpurun_out/r02a_bench_pk${PK}_gr${GR}.err").read()[-2000:])
PY
done; done
