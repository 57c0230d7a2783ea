#!/bin/bash
# usage: tools/run_scale.sh N tag   (under gpurun --gpus N)
N=$1; TAG=$2
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 1000 --warmup 3 > gpurun_out/${TAG}_scale_n$N.json 2> gpurun_out/${TAG}_scale_n$N.err
echo rc=$?
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/${TAG}_scale_n$N.json") if l.startswith("{")][-1]
print("value",round(d["value"],1),"ms/step",round(d["ms_per_step"],4),"e2e",round(d["e2e"]["value"],1),"one_ctx",d["one_context"]["ms_per_step_device"],d["one_context"]["ms_per_step_wall"])
for k in ("read_sharded","c4","c5"):
    v=d.get(k) or {}
    print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a in ("value","ms_per_step_wall","ms_per_batch","ms_per_step","error","balance_max_over_mean")})
print("parity", (d.get("sharded_parity") or {}).get("status"))
g=d.get("e2e_gz") or {}
print("e2e_gz", {k:(v.get("value"),v.get("all_runs_s")) for k,v in g.items() if isinstance(v,dict)}, g.get("inflate_threads"), g.get("error"))
PY
