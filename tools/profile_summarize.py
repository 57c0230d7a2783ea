"""Turn gpurun_out/{launches_TAG.csv, prof_TAG.ncu-rep, bench_TAG.json} into the tracked summaries under profiles/.
usage: python tools/profile_summarize.py TAG"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")

# ---- launch list
rows = [r for r in csv.reader(open(os.path.join(G, f"launches_{tag}.csv"))) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
d = collections.OrderedDict()
for r in rows[1:]:
    d.setdefault(r[ki].split("(")[0][-60:], []).append(float(r[vi].replace(",", "")))
with open(os.path.join(P, f"{tag}_launches.txt"), "w") as f:
    f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none, python bench.py --steps 2 --warmup 3 (tag {tag})\n")
    f.write("# per-launch device time (cold-cache, serialised under ncu: compare SHARES with bench.py's CUDA-event split)\n")
    # a step = the vk:: kernels that run in (nearly) every step; a context's FIRST k = 7 sample is counted by the flat-lane
    # kernel (one launch), the later ones by countt_kernel -- the one-off launch is listed but not summed
    most = max(len(v) for k, v in d.items() if "vk::" in k and "synth" not in k)
    step = {k: sum(v) / len(v) for k, v in d.items() if "synth" not in k and "vk::" in k and 2 * len(v) > most}
    tot = sum(step.values())
    for k, v in d.items():
        share = f"{100 * step[k] / tot:5.1f} % of a step" if k in step else ""
        f.write(f"{k:62s} launches={len(v):3d}  avg={sum(v) / len(v) / 1e3:9.1f} us  {share}\n")
    f.write(f"{'sum of one step':62s}               {tot / 1e3:9.1f} us\n")

# ---- full-set capture
rep = os.path.join(G, f"prof_{tag}.ncu-rep")
raw = os.path.join(G, f"prof_{tag}_raw.csv")
subprocess.run(f"ncu -i {rep} --page raw --csv > {raw} 2>/dev/null", shell=True, check=True)
out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), raw], capture_output=True, text=True).stdout
with open(os.path.join(P, f"{tag}_ncu_full.txt"), "w") as f:
    f.write(f"# ncu --set full --clock-control none --import-source on (tag {tag}); one block per profiled launch\n")
    f.write(out)
rows = list(csv.reader(open(raw)))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
def fnum(x):
    return float(x.replace(",", ""))
def to_bytes(val, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return fnum(val) * m.get(unit, 1)
for r in rows[2:]:
    if "countt_kernel<" in r[idx["Kernel Name"]] or "count_kernel<" in r[idx["Kernel Name"]]:
        rd = to_bytes(r[idx["dram__bytes_read.sum"]], rows[1][idx["dram__bytes_read.sum"]])
        wr = to_bytes(r[idx["dram__bytes_write.sum"]], rows[1][idx["dram__bytes_write.sum"]])
        kt = os.path.join(P, "kernel_traffic.json")          # per-kernel table that bench.py's roofline.traffic reads
        table = json.load(open(kt)) if os.path.exists(kt) else {}
        table["countt_kernel<16>" if "countt_kernel<" in r[idx["Kernel Name"]] else "count_kernel<7,smem>"] = {"dram_bytes_per_launch": rd + wr, "read": rd, "write": wr, "workload": "c2: 200 Mbp, k=7",
                                         "source": f"profiles/{tag}_ncu_full.txt (ncu --set full, tag {tag})"}
        json.dump(table, open(kt, "w"), indent=1)
        break
b = os.path.join(G, f"bench_{tag}.json")
if os.path.exists(b):
    line = [l for l in open(b) if l.startswith("{")][-1]
    open(os.path.join(P, f"{tag}_bench.json"), "w").write(line)
print(open(os.path.join(P, f"{tag}_launches.txt")).read())
