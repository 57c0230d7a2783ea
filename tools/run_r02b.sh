#!/bin/bash
# GPU run: full GPU suite, then the default bench with 1..4 samples in flight, graph on / off
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02b_pytest.log
tail -5 gpurun_out/r02b_pytest.log
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print({k:round(d[k],4) if isinstance(d[k],float) else d[k] for k in ("value","ms_per_step","ms_per_step_wall","gpu_launches","in_flight")}, "frac",round(d["roofline"]["frac"],4), "count_ms",round(d["roofline"]["kernel_ms"],4), {k:round(v,4) for k,v in d["kernel_ms_per_step"].items()}, "one_ctx dev/wall", round(d["one_context"]["ms_per_step_device"],4), round(d["one_context"]["ms_per_step_wall"],4), "e2e", round(d["e2e"]["value"],2))
except Exception as e:
    print("ERR", e); print(open(sys.argv[1].replace(".json",".err")).read()[-1500:])
PY
}
for GR in 1 0; do for T in 4 2 1; do
  VK_GRAPH=$GR timeout 300 python bench.py --steps 400 --warmup 3 --no-cpu-baseline --e2e-steps 4 --in-flight $T > gpurun_out/r02b_bench_gr${GR}_t${T}.json 2> gpurun_out/r02b_bench_gr${GR}_t${T}.err
  echo "gr=$GR T=$T rc=$?"; show gpurun_out/r02b_bench_gr${GR}_t${T}.json
done; done
VK_GRAPH=1 timeout 300 python bench.py --steps 400 --warmup 3 --no-cpu-baseline --e2e-steps 4 --in-flight 3 > gpurun_out/r02b_bench_gr1_t3.json 2> gpurun_out/r02b_bench_gr1_t3.err; show gpurun_out/r02b_bench_gr1_t3.json
VK_PACKED=1 timeout 300 python bench.py --steps 400 --warmup 3 --no-cpu-baseline --e2e-steps 4 > gpurun_out/r02b_bench_packed.json 2> gpurun_out/r02b_bench_packed.err; echo packed; show gpurun_out/r02b_bench_packed.json
