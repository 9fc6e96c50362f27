#!/usr/bin/env python
"""Diagnostic: per-kernel times of the contig-in-truth stage on cfg-2 shaped segments.  Run under
`ncu --metrics gpu__time_duration.sum -k regex:k_startpos --csv` to split the stage by kernel."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genomeassembler_dev_b200 import breakscore as B, synth, tables  # noqa: E402

n_seg = int(sys.argv[1]) if len(sys.argv) > 1 else 300
kmers, prob = tables.all_kmer_strings(), tables.normalised(tables.load_raw())
sc = B.BreakageScorer(0)
sc.set_table(kmers, prob)
b = synth.make_batch(n_seg, seed=1234)
args = (b.read_chars, None, b.read_len, b.contig_chars, b.contig_off, b.truth_chars, b.truth_off, b.seg_read_start, b.seg_contig_start)
sc.enable_timing(True)
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 3):
    sc.score_batch(*args, flags=B.DEFAULT_FLAGS)
    print({k: round(v, 3) for k, v in sc.last_timings().items() if v >= 0}, flush=True)
