#!/bin/bash
# A/B of two builds of the library on one box: the committed countt_kernel (libvk_prev.so) against the working tree
set -u
mkdir -p gpurun_out
cp varkoder_b200/libvarkoder_b200.so /tmp/cur.so
for V in cur prev cur prev; do
if [ $V = prev ]; then cp varkoder_b200/libvk_prev.so varkoder_b200/libvarkoder_b200.so; else cp /tmp/cur.so varkoder_b200/libvarkoder_b200.so; fi
VK_COUNT_LANES=2 timeout 600 python bench.py --steps 200 --warmup 3 --no-side-legs --no-cpu-baseline --e2e-steps 2 2>gpurun_out/r04w_bench.err | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][-1]
print('$V', round(d['value'],1), round(d['roofline']['frac'],4), {k:round(v,4) for k,v in d['kernel_ms_per_step'].items()})"
done
cp /tmp/cur.so varkoder_b200/libvarkoder_b200.so
