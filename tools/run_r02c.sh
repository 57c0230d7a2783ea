#!/bin/bash
# GPU run: sharded test (world 1), same-box A/B against the round-1 tree, bench
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sharded or variants or overflow" > gpurun_out/r02c_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02c_pytest.log
tail -5 gpurun_out/r02c_pytest.log
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print({k:round(d[k],4) if isinstance(d[k],float) else d[k] for k in ("value","ms_per_step","gpu_launches","in_flight")}, "frac",round(d["roofline"]["frac"],4), {k:round(v,4) for k,v in d["kernel_ms_per_step"].items()}, "one_ctx dev/wall", round(d["one_context"]["ms_per_step_device"],4), round(d["one_context"]["ms_per_step_wall"],4), "e2e", round(d["e2e"]["value"],2), "probe", d["roofline"].get("hbm_copy_probe_this_box_gbs"))
except Exception as e:
    print("ERR", e); print(open(sys.argv[1].replace(".json",".err")).read()[-1500:])
PY
}
timeout 300 python bench.py --steps 400 --warmup 3 --no-cpu-baseline --e2e-steps 4 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; echo now; show gpurun_out/r02c_bench.json
(cd _r01 && timeout 300 python bench.py --steps 400 --warmup 3 --no-cpu-baseline --e2e-steps 4 > ../gpurun_out/r02c_bench_r01tree.json 2> ../gpurun_out/r02c_bench_r01tree.err); echo r01-tree; show gpurun_out/r02c_bench_r01tree.json
