#!/bin/bash
# countt_kernel: parity of the variants, ncu full set + source page of the kernel
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "variants" > gpurun_out/r04b_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r04b_pytest.log
VK_COUNT_LANES=${LANES:-2} VK_GRAPH=0 timeout 600 ncu --set full --import-source on --clock-control none -k regex:countt -s 2 -c 1 -o gpurun_out/r04b_countt -f python tools/trace_step.py > gpurun_out/r04b_ncu.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/*.ncu-rep
