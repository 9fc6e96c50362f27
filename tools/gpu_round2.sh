#!/bin/bash
# Round-2 evidence in one gpurun call (run ON the GPU box from the repo root):
#   gpurun --timeout 1500 -- 'bash tools/gpu_round2.sh r02a'
# GPU suite, bench line, launch list, one --set full capture of EVERY kernel
# of one device-resident 1000-segment step.
set -u
tag=${1:-r02x}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "rc=$?" >> $out/${tag}_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench.err
rc=$?; echo "bench rc=$rc" >> $out/${tag}_bench.err
if [ $rc -ne 0 ]; then  # the study line alone if one of the extra measurements failed
  timeout 300 python bench.py --steps 5 --warmup 3 --skip-sharded > $out/${tag}_bench_1gpu_study_only.json 2>> $out/${tag}_bench.err
fi
if [ -n "${EXTRA_CMD:-}" ]; then bash -c "$EXTRA_CMD" > $out/${tag}_extra.log 2>&1; echo "rc=$?" >> $out/${tag}_extra.log; fi
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-study --scan-segments 0 --device-only > $out/${tag}_plain.log 2>&1 &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-study --scan-segments 0 --device-only > $out/${tag}_ncu_bench.log 2>&1
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-study --scan-segments 0 --device-only > $out/${tag}_plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:^k_|bs::k_" -s ${FULL_SKIP:-36} -c ${FULL_COUNT:-12} \
    -o $out/${tag}_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-study --scan-segments 0 --device-only > $out/${tag}_ncu_full.log 2>&1
ls -la $out | grep $tag
