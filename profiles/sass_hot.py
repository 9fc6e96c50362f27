"""Hot SASS instructions of one kernel in an ncu report (stall samples + executed counts).
    python profiles/sass_hot.py <report.ncu-rep> <kernel regex> [top]"""
import csv
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", f"regex:{pat}"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
data = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) == len(hdr):
        data.append(r)
i_s, i_ie, i_src = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
tot_s = sum(int(r[i_s]) for r in data)
tot_i = sum(int(r[i_ie]) for r in data)
print(f"{rows[0][1]}: {len(data)} SASS instructions, {tot_s} samples, {tot_i} warp instructions executed")
top = sorted(range(len(data)), key=lambda k: -int(data[k][i_s]))[:top_n]
for k in sorted(top):
    r = data[k]
    print(f"{k:5d} {int(r[i_s]):7d} ({100 * int(r[i_s]) / max(tot_s, 1):4.1f}%) {int(r[i_ie]):10d}  {r[i_src][:110]}")
