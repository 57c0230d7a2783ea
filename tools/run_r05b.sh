#!/bin/bash
# round 2, fourth session: the GPU suite against the library built with device-side assertions (make DEBUG=1, built in the
# container and shipped as libvarkoder_b200_debug.so), now that countt_kernel / countt9_kernel assert their staging addresses
set -u
mkdir -p gpurun_out
cp varkoder_b200/libvarkoder_b200.so /tmp/libvk_release.so
cp varkoder_b200/libvarkoder_b200_debug.so varkoder_b200/libvarkoder_b200.so
timeout 540 python -m pytest tests -m gpu -x -q > gpurun_out/r02d_debug_pytest.log 2>&1
echo "debug pytest rc=$?"; tail -4 gpurun_out/r02d_debug_pytest.log | cut -c1-300
cp /tmp/libvk_release.so varkoder_b200/libvarkoder_b200.so
