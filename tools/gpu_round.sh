#!/bin/bash
# One gpurun call's worth of evidence for a round (run ON the GPU box, from the repo root):
#   gpurun --timeout 900 -- 'bash tools/gpu_round.sh r02a "k_startpos_index|k_place_index"'
# 1. the GPU test-suite, 2. the bench line, 3. the ncu launch list of the same bench command (shares of the step),
# 4. one `--set full` capture of the named kernels on a single-chunk 1000-segment call (tools/diag_startpos_time.py
#    drives the whole pipeline once per iteration).  Everything lands in gpurun_out/<tag>_*; summarise afterwards with
#    `python profiles/summarize_ncu.py <tag> gpurun_out/<tag>_launches.csv gpurun_out/<tag>_full.ncu-rep`.
set -u
tag=${1:-rXX}
kernels=${2:-k_place_index}
out=gpurun_out
mkdir -p $out
timeout 400 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "rc=$?" >> $out/${tag}_pytest.log
timeout 200 python bench.py --steps 5 --warmup 3 > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench.err || exit 1
# candidates waiting for a GPU measurement ("Open"): same tests and bench with the switch on
BS_STARTPOS_BITMAP=1 timeout 200 python -m pytest tests -m gpu -x -q -k "startpos or cfg2_shape or cfg4 or cfg5 or edit_distance_edge or reference_vectors" > $out/${tag}_pytest_bitmap.log 2>&1; echo "rc=$?" >> $out/${tag}_pytest_bitmap.log
BS_STARTPOS_BITMAP=1 timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-study --scan-segments 0 > $out/${tag}_bench_bitmap.json 2>> $out/${tag}_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-study --scan-segments 0 > $out/${tag}_ncu_bench.log 2>&1
BS_CHUNK_KB=4000000 timeout 400 ncu --set full --clock-control none --import-source on -k "regex:$kernels" -s 2 -c 4 \
    -o $out/${tag}_full python tools/diag_startpos_time.py 1000 3 > $out/${tag}_ncu_full.log 2>&1
ls -la $out | grep $tag
