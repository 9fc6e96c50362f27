"""one bs_score_scaffolds call on the cfg-4 set after two warm-ups (profiling target)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from genomeassembler_dev_b200 import breakscore as B, synth, tables
kmers = tables.all_kmer_strings(); prob = tables.normalised(tables.load_raw())
seg = synth.make_scaffold_set(1400, n_scaffolds=int(os.environ.get("N_SCAF", "10000")))
sset = B.ScaffoldSet(seg.base_contigs, seg.part_start, seg.part_base, np.zeros(len(seg.part_base), np.int32))
sc = B.BreakageScorer(0)
sc.set_table(kmers, prob)
lens = sset.lengths()
for _ in range(3):
    res = sc.score_batch(seg.reads.reshape(-1), None, seg.reads.shape[1], np.zeros(1, np.uint8), np.concatenate([[0], np.cumsum(lens)]),
                         np.frombuffer(seg.truth, np.uint8), [0, len(seg.truth)], [0, len(seg.reads)], [0, len(sset)],
                         flags=B.WANT_KS | B.WANT_STARTPOS, scaffolds=sset)
print(int(res["kmer_breaks"].sum()))
