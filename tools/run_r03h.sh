#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "shard or expected_value or config4" > gpurun_out/r03h_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r03h_pytest.log
