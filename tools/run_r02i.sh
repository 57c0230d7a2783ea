#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "flood or pair or exact_16bit or u32_and_global or counts_bit_exact or variants or ladder_levels" > gpurun_out/r02i_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r02i_pytest.log
show() { python - "$1" <<'PY'
import json,sys
try:
    d=[json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")][-1]
    print({k:round(d[k],4) if isinstance(d[k],float) else d[k] for k in ("value","ms_per_step","in_flight")}, "frac",round(d["roofline"]["frac"],4), {k:round(v,4) for k,v in d["kernel_ms_per_step"].items()}, "one_ctx", round(d["one_context"]["ms_per_step_device"],4))
except Exception as e:
    print("ERR", e); print(open(sys.argv[1].replace(".json",".err")).read()[-1500:])
PY
}
for P in 0 1; do
VK_COUNT_PAIRS=$P timeout 300 python bench.py --steps 400 --warmup 3 --no-cpu-baseline --no-side-legs --e2e-steps 4 > gpurun_out/r02i_bench_pairs$P.json 2> gpurun_out/r02i_bench_pairs$P.err; echo pairs=$P; show gpurun_out/r02i_bench_pairs$P.json
done
VK_COUNT_PAIRS=1 VK_COUNT_FAST=0 timeout 300 python bench.py --steps 400 --warmup 3 --no-cpu-baseline --no-side-legs --e2e-steps 4 > gpurun_out/r02i_bench_pairs_exact.json 2> gpurun_out/r02i_bench_pairs_exact.err; echo pairs-exact; show gpurun_out/r02i_bench_pairs_exact.json
python tools/time_k.py 2>&1 | tail -12
VK_COUNT_FAST=0 python tools/time_k.py 2>&1 | tail -12
