"""Instruction regions of one kernel in an ncu report: runs of consecutive SASS instructions with (nearly) the same
execution count, with their share of the kernel's executed warp instructions and of its stall samples.
    python profiles/sass_regions.py <report.ncu-rep> <kernel regex> [top]"""
import csv
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 14
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", f"regex:{pat}"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
i_ie, i_src, i_s = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
data = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) == len(hdr) and r[i_ie].isdigit():
        data.append(r)
tot = sum(int(r[i_ie]) for r in data)
tot_s = sum(int(r[i_s]) for r in data)
print(f"{rows[0][1]}: {len(data)} SASS instructions, {tot} warp instructions executed, {tot_s} samples")
regions = []
for k, r in enumerate(data):
    c = int(r[i_ie])
    if regions and abs(regions[-1][2] - c) <= 0.02 * max(c, 1):
        regions[-1][1] = k
        regions[-1][3] += c
        regions[-1][4] += int(r[i_s])
    else:
        regions.append([k, k, c, c, int(r[i_s])])
regions.sort(key=lambda g: -g[3])
for g in regions[:top_n]:
    print(f"SASS {g[0]:5d}-{g[1]:5d} ({g[1] - g[0] + 1:4d} instr) executed ~{g[2]:>9d} x  {100 * g[3] / tot:5.1f}% of instructions  {100 * g[4] / max(tot_s, 1):5.1f}% of samples | {data[g[0]][i_src][:70]}")
