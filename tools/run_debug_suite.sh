#!/bin/bash
# GPU suite against a library built with device-side assertions (make DEBUG=1): every data-dependent address is checked
set -u
mkdir -p gpurun_out
cp varkoder_b200/libvarkoder_b200.so /tmp/libvk_release.so
make -s -C varkoder_b200/csrc clean >/dev/null 2>&1
make -s -C varkoder_b200/csrc DEBUG=1 > gpurun_out/debug_build.log 2>&1 || { tail -5 gpurun_out/debug_build.log; exit 1; }
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/debug_pytest.log 2>&1
echo "debug pytest rc=$?"; tail -4 gpurun_out/debug_pytest.log
cp /tmp/libvk_release.so varkoder_b200/libvarkoder_b200.so
