#!/bin/bash
# per-kernel device times of the cfg-4 / cfg-5 calls (ncu launch list restricted to a kernel regex), plus variants
set -u
tag=${1:-r02x}
regex=${2:-k_startpos}
out=gpurun_out
mkdir -p $out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:$regex" -c 4000 --csv --log-file $out/${tag}_cfg45_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-study --scan-segments 0 > $out/${tag}_cfg45_ncu.log 2>&1
echo "rc=$?" >> $out/${tag}_cfg45_ncu.log
