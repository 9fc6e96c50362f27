#!/bin/bash
# A/B of the fused-scoring threshold on ONE box (run ON the GPU box): device-only cfg-2 step, variants interleaved twice
set -u
out=gpurun_out; mkdir -p $out
f=$out/${1:-r02t}_fuse_ab.jsonl; : > $f
run() { echo "{\"variant\": \"$1\"}" >> $f; env $2 timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-study --scan-segments 0 --device-only 2>/dev/null | tail -1 >> $f; }
for rep in 1 2; do
  run unfused BS_FUSE_SCORE=0
  run min8k BS_FUSE_MIN_LEN=8192
  run min12k BS_FUSE_MIN_LEN=12288
  run min16k BS_FUSE_MIN_LEN=16384
  run min24k BS_FUSE_MIN_LEN=24576
done
python - "$f" <<'PY'
import json,sys
v=None
for l in open(sys.argv[1]):
    d=json.loads(l)
    if 'variant' in d: v=d['variant']; continue
    s=d['stage_ms_per_step']; print(v, round(d['ms_per_step'],3), 'score', round(s.get('score',0),3), 'pd', round(s['prob_dist_ks'],3))
PY
