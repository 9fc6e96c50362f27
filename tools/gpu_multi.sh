#!/bin/bash
# bench.py on N GPUs of one box (run ON the box: gpurun --gpus N -- 'bash tools/gpu_multi.sh <tag> N')
set -u
tag=${1:-r02x}
n=${2:-2}
out=gpurun_out
mkdir -p $out
if [ -n "${PRETEST:-}" ]; then timeout 600 python -m pytest tests -m gpu -x -q -k "$PRETEST" > $out/${tag}_pretest.log 2>&1; echo "rc=$?" >> $out/${tag}_pretest.log; fi
nvidia-smi topo -m > $out/${tag}_topo.txt 2>&1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $n --steps 5 --warmup 3 > $out/${tag}_bench_${n}gpu.json 2> $out/${tag}_bench_${n}gpu.err
echo "rc=$?" >> $out/${tag}_bench_${n}gpu.err
tail -c 600 $out/${tag}_bench_${n}gpu.err
