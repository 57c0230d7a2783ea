#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "beyond_4_gib or epochs or fire_and_forget" > gpurun_out/r04aa_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r04aa_pytest.log
python __graft_entry__.py smoke 2>&1 | tail -2
