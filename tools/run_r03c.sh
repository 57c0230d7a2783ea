#!/bin/bash
# config 4 (96 small samples): samples in flight x cap on the count CTAs of a small sample; per-launch trace of a 30 Mbp sample
set -u
mkdir -p gpurun_out
for T in 4 8 12; do for RPC in 0 1500 3000; do
  VK_COUNT_READS_PER_CTA=$RPC python bench.py --workload c4 --in-flight $T --steps 3 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][-1]
print('T=$T RPC=$RPC', round(d['value'],1), round(d['ms_per_step'],3))"
done; done
VK_TRACE_EACH=1 VK_N=30000000 VK_NOMAX=1 VK_MAP=varKode python tools/trace_step.py 2>&1 | grep "vk trace" | tail -12
