"""Per SOURCE LINE totals of one kernel in an ncu report: executed warp instructions and stall samples, the SASS of the
report matched by position with `nvdisasm -g` of the cubin inside the library that was profiled (built with -lineinfo).
    python profiles/sass_lines.py <report.ncu-rep> <kernel regex> <mangled kernel name> [library.so] [top]"""
import collections
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

rep, pat, mangled = sys.argv[1], sys.argv[2], sys.argv[3]
lib = sys.argv[4] if len(sys.argv) > 4 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "genomeassembler_dev_b200", "libbreakscore.so")
top_n = int(sys.argv[5]) if len(sys.argv) > 5 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", f"regex:{pat}"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
i_ie, i_s = hdr.index("Instructions Executed"), hdr.index("# Samples")
data = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) == len(hdr) and r[i_ie].isdigit():
        data.append((int(r[i_ie]), int(r[i_s])))
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=td, capture_output=True)
    lines = []
    for cub in sorted(glob.glob(os.path.join(td, "*.cubin"))):
        out = subprocess.run(["nvdisasm", "-g", cub], capture_output=True, text=True).stdout.splitlines()
        inside, cur = False, ("?", 0)
        for ln in out:
            if ln.startswith("\t.section") or ln.startswith(".section"):
                inside = (".text." + mangled) in ln
                continue
            if not inside:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", ln):
                lines.append(cur)
        if lines:
            break
n = min(len(lines), len(data))
print(f"{rows[0][1]}: {len(data)} SASS instructions in the report, {len(lines)} in the cubin")
agg = collections.defaultdict(lambda: [0, 0])
for k in range(n):
    agg[lines[k]][0] += data[k][0]
    agg[lines[k]][1] += data[k][1]
ti, ts = sum(d[0] for d in data), sum(d[1] for d in data)
for (f, l), (ie, sm) in sorted(agg.items(), key=lambda kv: -kv[1][int(os.environ.get("SORT_COL", "1"))])[:top_n]:
    print(f"{f}:{l:<5d} {100 * ie / max(ti, 1):5.1f}% of instructions  {100 * sm / max(ts, 1):5.1f}% of samples")
