#!/bin/bash
# Device-resident step of the 1000-segment study with several builds of the library (tuning variants made with
# `make OUT=... EXTRA_NVFLAGS=-D...`): one JSON line per variant in gpurun_out/<tag>_variants.jsonl
set -u
tag=${1:-r02x}
shift
out=gpurun_out
mkdir -p $out
: > $out/${tag}_variants.jsonl
for lib in "$@"; do
  echo "{\"variant\": \"$lib\"}" >> $out/${tag}_variants.jsonl
  BREAKSCORE_LIB=$PWD/genomeassembler_dev_b200/$lib timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --device-only >> $out/${tag}_variants.jsonl 2>> $out/${tag}_variants.err
done
