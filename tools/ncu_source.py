"""ncu --page source --csv: per-instruction executed counts and stall samples.  usage: ncu_source.py file.csv [pattern ...]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]
ix = {n: i for i, n in enumerate(h)}
pats = sys.argv[2:]
tot = samp = 0
base = None
for r in rows[2:]:
    if len(r) < 10:
        continue
    a = int(r[ix['Address']], 16)
    base = a if base is None else base
    n = int(r[ix['Instructions Executed']]); s = int(r[ix['# Samples']])
    tot += n; samp += s
    src = r[ix['Source']].strip()
    if not pats or any(p in src for p in pats):
        print(f"{a - base:5x} {n:10d} {s:6d} thr {r[ix['Avg. Threads Executed']]:>5s}  {src[:90]}")
print("total inst", tot, "samples", samp)
