#!/bin/bash
# round 2, third session, final records: full GPU suite, full bench line (side legs), reference arm, c3, ncu launch list + full capture
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r02c_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02c_pytest.log
python bench.py --steps 1000 --warmup 3 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=[json.loads(l) for l in open("gpurun_out/r02c_bench.json") if l.startswith("{")][-1]
print({k:round(d[k],4) if isinstance(d[k],float) else d[k] for k in ("value","ms_per_step","in_flight","gpu_launches")}, "frac",round(d["roofline"]["frac"],4), d["roofline"]["kernel"], {k:round(v,4) for k,v in d["kernel_ms_per_step"].items()}, "one_ctx", d["one_context"]["ms_per_step_device"], "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"])
for k in ("c4","c5"): print(k, d[k]["value"])
print(d["e2e_gz"]["one_sample"]["value"], d["e2e_gz"]["batch"]["value"], "level_bases", d["level_bases"])
PY
python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-300
python bench.py --workload c3 --steps 50 --warmup 3 --no-cpu-baseline --no-side-legs > gpurun_out/r02c_bench_c3.json 2> gpurun_out/r02c_bench_c3.err; python - <<'PY'
import json
d=[json.loads(l) for l in open("gpurun_out/r02c_bench_c3.json") if l.startswith("{")][-1]
print("c3", round(d["value"],1), round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["kernel_ms_per_step"].items()}, d["roofline"]["frac"])
PY
bash tools/profile_gpu.sh r02c
ncu -i gpurun_out/prof_r02c.ncu-rep --page raw --csv > gpurun_out/prof_r02c_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/prof_r02c_raw.csv > gpurun_out/r02c_ncu_full.txt 2>&1; wc -l gpurun_out/r02c_ncu_full.txt
