#!/usr/bin/env python
"""profiles/<tag>_ncu_full.txt (summarize_ncu.py's digest of one `ncu --set full` capture of a device-resident
1000-segment step) -> profiles/traffic_cfg2.json: dram__bytes_read.sum + dram__bytes_write.sum per launch and kernel,
which bench.py puts beside the compulsory bytes of every stage (roofline.traffic).

    python tools/traffic_from_ncu.py r02d [segments]
"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNITS = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def to_bytes(text):
    val, unit = text.split()[:2]
    return float(val.replace(",", "")) * UNITS[unit]


def main():
    tag = sys.argv[1]
    segments = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    path = os.path.join(ROOT, "profiles", f"{tag}_ncu_full.txt")
    out = {"segments": segments,
           "source": f"profiles/{tag}_ncu_full.txt (ncu --set full --clock-control none --import-source on, bench.py --steps 1 --warmup 3 "
                     f"--device-only: every kernel of the timed device-resident step; k_pack_seqs = contigs + truths launches summed)",
           "unit": "bytes per step: dram__bytes_read.sum + dram__bytes_write.sum"}
    for block in open(path).read().split("-----\n")[1:]:
        d = {}
        for line in block.strip().splitlines():
            d[line[:70].strip()] = line[70:].strip()
        name = re.sub(r"^void ", "", d["Kernel Name"])
        name = re.sub(r"[<(].*$", "", name)
        out[name] = out.get(name, 0.0) + to_bytes(d["dram__bytes_read.sum"]) + to_bytes(d["dram__bytes_write.sum"])
    with open(os.path.join(ROOT, "profiles", "traffic_cfg2.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
