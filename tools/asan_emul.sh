#!/bin/bash
# Memory-safety check of the kernel sources without a GPU: compute-sanitizer is closed on the
# GPU pool, so the same .cu/.cuh files are compiled for the CPU emulation (tests/emul/cuda_emul.h)
# with AddressSanitizer + UBSan and driven through every placement mode, the edge inputs and the
# edit distance (tools/asan_emul_cases.py).  Test tooling only.
set -e
cd "$(dirname "$0")/.."
mkdir -p /tmp/asan
g++ -O1 -g -std=c++20 -DBS_CPU_EMUL -fPIC -shared -pthread -fsanitize=address,undefined -fno-omit-frame-pointer \
    -I tests/emul -I genomeassembler_dev_b200/csrc -x c++ genomeassembler_dev_b200/csrc/bs_api.cu \
    genomeassembler_dev_b200/csrc/bs_assemble.cpp -o /tmp/asan/libbreakscore_emul.so
ASAN_OPTIONS=detect_leaks=0 LD_PRELOAD=$(g++ -print-file-name=libasan.so):$(g++ -print-file-name=libubsan.so) \
    python tools/asan_emul_cases.py
