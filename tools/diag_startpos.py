#!/usr/bin/env python
"""Diagnostic: contig-in-truth offsets of a cfg-5 shaped call (one 1 Mb truth, 1e5 reads, 1001 contigs)
against the known substring offsets, repeated with the dense and the hashed placement scratch."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genomeassembler_dev_b200 import breakscore as B, synth, tables  # noqa: E402

kmers, prob = tables.all_kmer_strings(), tables.normalised(tables.load_raw())
sc = B.BreakageScorer(0, sys.argv[4] if len(sys.argv) > 4 and sys.argv[4] != "-" else None)
sc.set_table(kmers, prob)
rng = np.random.default_rng(506)
L, N, Cn, r = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), 150
truth = synth.codes_to_ascii(synth.random_truth_codes(rng, L))
starts = rng.integers(0, L - r, size=N)
reads = truth[starts[:, None] + np.arange(r)[None, :]]
cstart = np.sort(rng.integers(0, L - 3000, size=Cn))
clen = np.append(rng.integers(200, 2000, size=Cn), 60000)
contigs = [truth[a:a + b].tobytes() for a, b in zip(cstart, clen[:-1])]
contigs.append(truth[5000:65000].tobytes())
exp = np.append(cstart, 5000).astype(np.int32)
for it in range(int(sys.argv[5]) if len(sys.argv) > 5 else 6):
    res = sc.score(contigs, reads, truth.tobytes(), flags=B.WANT_STARTPOS if it % 2 else B.DEFAULT_FLAGS)
    want = exp * (res["kmer_breaks"] > 0)
    got = res["path_prob_dist_startpos"]
    bad = np.nonzero(got != want)[0]
    print("call", it, "n_bad", len(bad), "placed_contigs", int((res["kmer_breaks"] > 0).sum()),
          "idx", bad[:6].tolist(), "got", got[bad[:6]].tolist(), flush=True)
