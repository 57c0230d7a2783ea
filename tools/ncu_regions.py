"""Group the stall samples of one kernel (ncu --page source --csv) by SASS regions that end at a barrier / branch.
usage: python tools/ncu_regions.py report.ncu-rep kernel_regex [min_share]"""
import csv
import io
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = []
        blocks.append(cur)
        continue
    if cur is not None:
        cur.append(r)
b = blocks[0]
h = b[0]
isrc, ismp, ie = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
tot = sum(int(r[ismp]) for r in b[1:] if len(r) > ismp and r[ismp].isdigit())
print("total samples", tot, "SASS lines", len(b) - 1)
acc = inst = 0
for i, r in enumerate(b[1:]):
    if len(r) <= ismp or not r[ismp].isdigit():
        continue
    acc += int(r[ismp])
    inst += int(r[ie])
    s = r[isrc]
    if any(x in s for x in ("BAR.", "UCGABAR", "EXIT", "BRA", "WARPSYNC")):
        if acc:
            print(f"{i:5d} {s[:50]:50s} samples {acc:6d} ({100 * acc / tot:5.1f} %)  warp-instr {inst}")
        acc = inst = 0
