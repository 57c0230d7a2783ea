#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "epochs or variants" > gpurun_out/r04x_pytest.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r04x_pytest.log
bash tools/run_r04w.sh
VK_N=15000000000 python tools/diag_big.py 2>&1 | tail -2
