#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ladder or shard" > gpurun_out/r03f_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r03f_pytest.log
python bench.py --steps 400 --warmup 3 --no-side-legs --no-cpu-baseline --e2e-steps 2 > gpurun_out/r03f_bench.json 2> gpurun_out/r03f_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=[json.loads(l) for l in open("gpurun_out/r03f_bench.json") if l.startswith("{")][-1]
print({k:round(d[k],4) if isinstance(d[k],float) else d[k] for k in ("value","ms_per_step","in_flight","gpu_launches")}, "frac",round(d["roofline"]["frac"],4), {k:round(v,4) for k,v in d["kernel_ms_per_step"].items()}, "one_ctx", d["one_context"]["ms_per_step_device"])
PY
VK_TRACE_EACH=1 python tools/trace_step.py 2>&1 | grep "vk trace" | tail -11
