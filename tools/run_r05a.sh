#!/bin/bash
# round 2, fourth session, records of the final tree (17 GPU-minutes left): full GPU suite, config 3 with countt9_kernel,
# ncu --set full + launch list of the k = 9 step, then the driver's bench command
set -u
mkdir -p gpurun_out
timeout 450 python -m pytest tests -m gpu -x -q > gpurun_out/r02d_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02d_pytest.log | cut -c1-300
timeout 200 python bench.py --workload c3 --steps 50 --warmup 3 --no-cpu-baseline --no-side-legs > gpurun_out/r02d_bench_c3.json 2> gpurun_out/r02d_bench_c3.err; echo "c3 rc=$?"
python - <<'PY'
import json
d=[json.loads(l) for l in open("gpurun_out/r02d_bench_c3.json") if l.startswith("{")][-1]
print("c3", round(d["value"],1), round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["kernel_ms_per_step"].items()}, d["roofline"]["frac"], d["roofline"]["kernel"])
PY
timeout 150 python tools/time_c3.py > gpurun_out/r02d_time_c3.log 2>&1; echo "time_c3 rc=$?"; tail -1 gpurun_out/r02d_time_c3.log | cut -c1-400
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"countt9_kernel" -s 3 -c 1 -o gpurun_out/prof_r02d_countt9 -f python tools/time_c3.py > gpurun_out/r02d_ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/r02d_ncu_full.log
ncu -i gpurun_out/prof_r02d_countt9.ncu-rep --page raw --csv > gpurun_out/prof_r02d_countt9_raw.csv 2>/dev/null
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r02d_c3.csv python tools/time_c3.py > gpurun_out/r02d_ncu_launches.log 2>&1
echo "launch list rc=$?"
timeout 400 python bench.py > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=[json.loads(l) for l in open("gpurun_out/r02d_bench.json") if l.startswith("{")][-1]
print({k:round(d[k],4) if isinstance(d[k],float) else d[k] for k in ("value","ms_per_step","in_flight","gpu_launches")}, "frac",round(d["roofline"]["frac"],4), d["roofline"]["kernel"], "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"])
for k in ("c4","c5"): print(k, d[k]["value"])
PY
