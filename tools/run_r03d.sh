#!/bin/bash
# full GPU suite + bench after the calibrated-threshold change
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r03d_pytest.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/r03d_pytest.log
python bench.py --steps 500 --warmup 3 --no-side-legs > gpurun_out/r03d_bench.json 2> gpurun_out/r03d_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=[json.loads(l) for l in open("gpurun_out/r03d_bench.json") if l.startswith("{")][-1]
print({k:round(d[k],4) if isinstance(d[k],float) else d[k] for k in ("value","ms_per_step","in_flight","gpu_launches")}, "frac",round(d["roofline"]["frac"],4), {k:round(v,4) for k,v in d["kernel_ms_per_step"].items()}, "one_ctx", d["one_context"]["ms_per_step_device"], "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"])
print("level_bases", d.get("level_bases"))
PY
