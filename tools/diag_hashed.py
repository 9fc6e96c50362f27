#!/usr/bin/env python
"""Diagnostic: where the wall time of a cfg-5 shaped call goes with the hashed placement scratch vs dense rows
(BS_TRACE=1 prints the host time of every stage's queueing code)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genomeassembler_dev_b200 import breakscore as B, synth, tables  # noqa: E402

kmers, prob = tables.all_kmer_strings(), tables.normalised(tables.load_raw())
sc = B.BreakageScorer(0)
sc.set_table(kmers, prob)
rng = np.random.default_rng(1500)
L, N, Cn, r = 10_000_000, 2_000_000, 10_000, 150
truth = synth.codes_to_ascii(synth.random_truth_codes(rng, L))
starts = rng.integers(0, L - r, size=N)
reads = truth[starts[:, None] + np.arange(r)[None, :]]
cstart = np.sort(rng.integers(0, L - 3000, size=Cn))
clen = rng.integers(200, 2000, size=Cn)
seg = synth.Segment(truth.tobytes(), reads, [truth[a:a + b].tobytes() for a, b in zip(cstart, clen)])
ct, ct_off = B.flatten(seg.contigs)
rd = np.ascontiguousarray(seg.reads).reshape(-1)
tr, tr_off = B.flatten([seg.truth])
args = (rd, None, seg.reads.shape[1], ct, ct_off, tr, tr_off, [0, seg.reads.shape[0]], [0, len(seg.contigs)])
for mode in ("hashed", "dense", "hashed"):
    if mode == "dense":
        os.environ["BS_PLACE_SCRATCH_MB"] = "65536"
    else:
        os.environ.pop("BS_PLACE_SCRATCH_MB", None)
    for it in range(5):
        if it == 4:
            os.environ["BS_TRACE"] = "1"
        t0 = time.perf_counter()
        sc.score_batch(*args, flags=B.DEFAULT_FLAGS)
        print(mode, it, "wall_ms", round(1e3 * (time.perf_counter() - t0), 2), flush=True)
        os.environ.pop("BS_TRACE", None)
