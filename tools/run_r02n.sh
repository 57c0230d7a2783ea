#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "variants or pair or overflow or long_reads or ladder_levels or edge" > gpurun_out/r02n_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02n_pytest.log
show() { python - "$1" <<'PY'
import json,sys
try:
    d=[json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")][-1]
    print({k:round(d[k],4) if isinstance(d[k],float) else d[k] for k in ("value","ms_per_step","in_flight")}, "frac",round(d["roofline"]["frac"],4), {k:round(v,4) for k,v in d["kernel_ms_per_step"].items()}, "one_ctx", round(d["one_context"]["ms_per_step_device"],4))
except Exception as e:
    print("ERR", e); print(open(sys.argv[1].replace(".json",".err")).read()[-1500:])
PY
}
run() { env "$@" timeout 300 python bench.py --steps 400 --warmup 3 --no-cpu-baseline --no-side-legs --e2e-steps 4 > gpurun_out/r02n_b.json 2> gpurun_out/r02n_b.err; echo "$@"; show gpurun_out/r02n_b.json; }
run VK_CHUNKS=1 VK_COUNT_PAIRS=0
run VK_CHUNKS=1 VK_COUNT_PAIRS=1
run VK_CHUNKS=0 VK_COUNT_PAIRS=0
