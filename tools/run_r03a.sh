#!/bin/bash
# round 3 session A: the lane-per-read pair kernel (VK_COUNT_LANES=1): variants test, A/B bench on one box
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "step_variants" > gpurun_out/r03a_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r03a_pytest.log
for L in 0 1; do
  VK_COUNT_LANES=$L python bench.py --steps 300 --warmup 3 --no-cpu-baseline --no-side-legs > gpurun_out/r03a_bench_$L.json 2> gpurun_out/r03a_bench_$L.err; echo "bench $L rc=$?"
  python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/r03a_bench_$L.json") if l.startswith("{")][-1]
print("lanes=$L", {k:round(d[k],4) if isinstance(d[k],float) else d[k] for k in ("value","ms_per_step")}, "frac",round(d["roofline"]["frac"],4), {k:round(v,4) for k,v in d["kernel_ms_per_step"].items()}, "one_ctx", d["one_context"]["ms_per_step_device"])
PY
done
