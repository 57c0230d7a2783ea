#!/bin/bash
# A/B of the count kernel only: VK_COUNT_LANES=0/1, short bench (kernel_ms_per_step.count), optional pytest first
set -u
mkdir -p gpurun_out
if [ "${1:-}" = "test" ]; then
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "step_variants and 1-0-0-1" > gpurun_out/r03b_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r03b_pytest.log
fi
for L in ${LANES:-1}; do
  VK_COUNT_LANES=$L python bench.py --steps 200 --warmup 3 --no-cpu-baseline --no-side-legs --e2e-steps 2 > gpurun_out/r03b_bench_$L.json 2> gpurun_out/r03b_bench_$L.err; echo "bench $L rc=$?"
  python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/r03b_bench_$L.json") if l.startswith("{")][-1]
print("lanes=$L", {k:round(d[k],4) if isinstance(d[k],float) else d[k] for k in ("value","ms_per_step")}, "frac",round(d["roofline"]["frac"],4), {k:round(v,4) for k,v in d["kernel_ms_per_step"].items()}, "one_ctx", d["one_context"]["ms_per_step_device"])
PY
done
