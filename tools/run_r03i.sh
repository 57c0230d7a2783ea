#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "very_long or shard or ladder or overflow or edge" > gpurun_out/r03i_pytest.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/r03i_pytest.log
