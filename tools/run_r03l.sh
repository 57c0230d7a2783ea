#!/bin/bash
# per-kernel device times of a 30 Mbp sample (config-4 size), ncu launch list
set -u
mkdir -p gpurun_out
VK_GRAPH=0 VK_N=30000000 VK_NOMAX=1 VK_MAP=varKode ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_small.csv python tools/trace_step.py > gpurun_out/ncu_small.log 2>&1
echo rc=$?
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open("gpurun_out/launches_small.csv")) if len(r)>10]
h=rows[0]; ki,vi=h.index("Kernel Name"),h.index("Metric Value")
d=collections.OrderedDict()
for r in rows[1:]: d.setdefault(r[ki].split("(")[0][-40:],[]).append(float(r[vi].replace(",","")))
tot=0
for k,v in d.items():
    a=sum(v[-3:])/len(v[-3:])/1e3
    if "synth" not in k and "Fill" not in k: tot+=a
    print(f"{k:42s} n={len(v):3d} last3 avg {a:8.1f} us")
print("sum", round(tot,1))
PY
