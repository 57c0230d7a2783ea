#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02l_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r02l_pytest.log
show() { python - "$1" <<'PY'
import json,sys
try:
    d=[json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")][-1]
    print({k:round(d[k],4) if isinstance(d[k],float) else d[k] for k in ("value","ms_per_step","in_flight")}, "frac",round(d["roofline"]["frac"],4), {k:round(v,4) for k,v in d["kernel_ms_per_step"].items()}, "one_ctx", round(d["one_context"]["ms_per_step_device"],4))
except Exception as e:
    print("ERR", e); print(open(sys.argv[1].replace(".json",".err")).read()[-1500:])
PY
}
for C in 1 0 P; do
if [ $C = P ]; then export VK_COUNT_PAIRS=1; C=1; fi; VK_CHUNKS=$C timeout 300 python bench.py --steps 400 --warmup 3 --no-cpu-baseline --no-side-legs --e2e-steps 4 > gpurun_out/r02l_bench_chunks${C}_p${VK_COUNT_PAIRS:-0}.json 2> gpurun_out/r02l_bench_chunks${C}_p${VK_COUNT_PAIRS:-0}.err; echo chunks=$C; show gpurun_out/r02l_bench_chunks${C}_p${VK_COUNT_PAIRS:-0}.json
done
VK_COUNT_PAIRS=1 VK_TRACE_EACH=1 python tools/trace_step.py 2> gpurun_out/r02l_trace.log; tail -12 gpurun_out/r02l_trace.log | cut -c1-110
