"""Turns the scratch ncu outputs under gpurun_out/ into the small committed summaries here.

    python profiles/summarize_ncu.py <tag> [launches.csv] [report.ncu-rep]

writes profiles/<tag>_launch_shares.txt (per-kernel share of the step from the
`--metrics gpu__time_duration.sum` pass) and profiles/<tag>_ncu_full.txt (selected metrics of the
`--set full` capture, per launch)."""
import collections
import csv
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
WANT = [
    "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "launch__waves_per_multiprocessor",
]


def shares(path, out):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        k = row["Kernel Name"]
        v = float(row["Metric Value"].replace(",", ""))
        a = agg.setdefault(k, [0, 0.0, row["Metric Unit"]])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"# per-kernel device time from `ncu --metrics gpu__time_duration.sum --clock-control none` ({os.path.basename(path)})\n")
        f.write("# cold-cache, serialised launches: compare SHARES with bench.py's stage_ms_per_step, not absolutes\n")
        for k, a in agg.items():
            f.write(f"{k[:60]:62s} launches={a[0]:4d} total={a[1]:14.1f} {a[2]} share={a[1] / tot:.3f}\n")


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [(w, hdr.index(w)) for w in WANT if w in hdr]
    with open(out, "w") as f:
        f.write(f"# selected metrics of `ncu --set full --clock-control none --import-source on` ({os.path.basename(rep)}), one block per launch\n")
        for r in rows[2:]:
            f.write("-----\n")
            for w, i in idx:
                f.write(f"{w:70s} {r[i]} {units[i]}\n")


if __name__ == "__main__":
    tag = sys.argv[1]
    if len(sys.argv) > 2 and sys.argv[2] != "-":
        shares(sys.argv[2], os.path.join(HERE, f"{tag}_launch_shares.txt"))
    if len(sys.argv) > 3:
        full(sys.argv[3], os.path.join(HERE, f"{tag}_ncu_full.txt"))
