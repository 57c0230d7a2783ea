#!/bin/bash
mkdir -p gpurun_out
VK_N=15000000000 python tools/diag_big.py 2>&1 | tail -5
VK_N=15000000000 VK_K=8 python tools/diag_big.py 2>&1 | tail -4
