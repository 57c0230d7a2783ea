#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02p_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02p_pytest.log
python bench.py --steps 1000 --warmup 3 > gpurun_out/r02p_bench.json 2> gpurun_out/r02p_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=[json.loads(l) for l in open("gpurun_out/r02p_bench.json") if l.startswith("{")][-1]
print({k:round(d[k],4) if isinstance(d[k],float) else d[k] for k in ("value","ms_per_step","in_flight","gpu_launches")}, "frac",round(d["roofline"]["frac"],4), {k:round(v,4) for k,v in d["kernel_ms_per_step"].items()}, "one_ctx", d["one_context"]["ms_per_step_device"], "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"])
for k in ("c4","c5"): print(k, d[k]["value"])
print(d["e2e_gz"]["one_sample"]["value"], d["e2e_gz"]["batch"]["value"])
PY
python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-400
VK_N=40000000 python tools/time_k.py 7 | tail -1
VK_N=200000000 python tools/time_k.py 7 5 6 8 | tail -4
python bench.py --workload c3 --steps 50 --warmup 3 --no-cpu-baseline --no-side-legs > gpurun_out/r02p_bench_c3.json 2> gpurun_out/r02p_bench_c3.err; python - <<'PY'
import json
d=[json.loads(l) for l in open("gpurun_out/r02p_bench_c3.json") if l.startswith("{")][-1]
print("c3", round(d["value"],1), round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["kernel_ms_per_step"].items()}, d["roofline"])
PY
