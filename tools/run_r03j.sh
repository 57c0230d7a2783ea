#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "very_long or shard or ladder or overflow or variants" > gpurun_out/r03j_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r03j_pytest.log
python bench.py --steps 400 --warmup 3 --no-side-legs --no-cpu-baseline --e2e-steps 2 > gpurun_out/r03j_bench.json 2> gpurun_out/r03j_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=[json.loads(l) for l in open("gpurun_out/r03j_bench.json") if l.startswith("{")][-1]
print({k:round(d[k],4) if isinstance(d[k],float) else d[k] for k in ("value","ms_per_step","in_flight","gpu_launches")}, "frac",round(d["roofline"]["frac"],4), {k:round(v,4) for k,v in d["kernel_ms_per_step"].items()}, "one_ctx", d["one_context"]["ms_per_step_device"])
PY
