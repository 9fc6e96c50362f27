"""Case lists and the checker shared by the CPU-emulation tests and the GPU parity tests."""
import numpy as np

from conftest import assert_matches_reference, assert_same_as_oracle
from genomeassembler_dev_b200 import breakscore as B
from genomeassembler_dev_b200 import synth

FULL = B.DEFAULT_FLAGS | B.WANT_HIST | B.WANT_POS

# (seed, L, read_len, coverage, n_contigs, n_gap_scaffolds): small enough for the oracle in seconds
SMALL = [
    (21, 3000, 40, 10, 6, 1),
    (22, 2500, 100, 8, 5, 0),
    (23, 2000, 12, 10, 8, 1),
    (24, 2600, 150, 10, 4, 1),
    (25, 1500, 33, 10, 5, 0),
    (26, 4000, 300, 10, 3, 0),
]
MEDIUM = [
    (31, 50000, 100, 30, 16, 1),   # cfg-1 shape
    (32, 50000, 150, 30, 24, 2),   # one cfg-2 segment
    (33, 50000, 40, 40, 16, 0),    # the reference's own grid (script 00)
    (34, 50000, 12, 40, 30, 0),    # shortest reads of script 00: leftmost rule under many repeats
    (35, 50000, 300, 30, 8, 0),
    (36, 20000, 64, 30, 12, 1),    # read length a multiple of 32
    (37, 20000, 65, 30, 12, 1),
    # cfg-3: read-length sweep 50-300 bp; candidate sets get shorter and more numerous with larger k
    (38, 50000, 50, 30, 120, 0),
    (39, 50000, 75, 30, 80, 1),
    (40, 50000, 200, 30, 40, 0),
    (41, 50000, 250, 30, 25, 1),
    (42, 50000, 14, 40, 60, 0),    # script 00 grid, second row
    (43, 50000, 25, 40, 40, 0),
]


def check_segment(scorer, oracle, kmers, prob, seg, kmer=8, flags=FULL, truth_prob=None, reads=None):
    reads = seg.read_list if reads is None else reads
    scorer.set_table(kmers, prob, truth_prob)
    got = scorer.score(seg.contigs, reads, seg.truth, kmer=kmer, flags=flags)
    want = oracle.oracle_calc_breakscore(seg.contigs, reads, seg.truth, kmer, kmers, prob, truth_prob=truth_prob,
                                         want_pos=bool(flags & B.WANT_POS), want_hist=bool(flags & B.WANT_HIST),
                                         want_lev=bool(flags & B.WANT_LEV))
    assert_same_as_oracle(got, want, check_pos=bool(flags & B.WANT_POS), check_hist=bool(flags & B.WANT_HIST))
    return got, want


def check_reference_vector(scorer, case, kmers, table_set):
    scorer.set_table(kmers, table_set[case["table"]])
    got = scorer.score(case["path"], case["reads"], case["truth"], kmer=case["kmer"], flags=B.DEFAULT_FLAGS)
    assert_matches_reference(got, case["expected"])
    if case["table"] in ("rowid", "rowsq"):
        assert np.array_equal(got["bp_score"], case["expected"]["bp_score"])


def make(seed, L, r, cov, nc, gaps, mut=0.25):
    return synth.make_segment(seed, length=L, read_len=r, coverage=cov, n_contigs=nc, mut_frac=mut, n_gap_scaffolds=gaps)


def edge_inputs():
    """(name, contigs, reads, truth, kmer): ragged and degenerate inputs"""
    c = b"ACGTTGCAAGGCTTACCGATAGGA"
    rep = (b"ACGGTCA" * 40)
    return [
        ("empty_read_set", [c, c[2:20]], [], c, 8),
        ("empty_read_string", [c], [b"", c[3:9]], c, 8),
        ("read_longer_than_contig", [c[:10]], [c[:12], c[2:8]], c, 8),
        ("contig_shorter_than_kmer", [b"ACGT", c], [b"CG", c[5:12]], c, 8),
        ("empty_contig", [b"", c], [c[5:12]], c, 8),
        ("non_acgt_reads", [b"ACGNTGCAAGGCTTNNCGATAGGA", c], [b"GNTGC", b"TNNCG", b"NN", b"acgt", b"GGCTT"], c, 8),
        ("lowercase_contig", [b"ACGTTGCAaggcTTACCGATAGGA"], [b"GCAagg", b"GCAAGG", b"TTACC"], c, 8),
        ("ragged_reads", [c, rep], [c[i:i + 5 + i] for i in range(0, 10)] + [rep[3:50], rep[1:9]], c + rep, 8),
        ("kmer4", [c, rep], [c[i:i + 8] for i in range(0, 12)], c, 4),
        ("kmer6", [c, rep], [c[i:i + 8] for i in range(0, 12)], c, 6),
        ("kmer2", [c], [c[i:i + 8] for i in range(0, 12)], c, 2),
        ("kmer10_longer_than_any_row", [c, rep], [c[i:i + 8] for i in range(0, 12)], c, 10),
        ("kmer40_longer_than_a_word", [c, rep], [c[i:i + 8] for i in range(0, 12)] + [rep[5:60]], c + rep, 40),
        ("kmer3_not_in_table", [c], [c[i:i + 8] for i in range(0, 12)], c, 3),
        ("contig_longer_than_truth", [c + c], [c[4:14]], c, 8),
        ("truth_shorter_than_kmer", [c], [c[4:14]], b"ACG", 8),
        ("all_same_base", [b"A" * 200], [b"A" * 20, b"A" * 33, b"A" * 64, b"AAAC"], b"A" * 300, 8),
        ("n_contig_in_n_truth", [rep[40:90] + b"NNNN" + rep[3:40], rep[50:90] + b"NNNN" + rep[3:10] + b"T"],
         [rep[45:70], rep[60:80]], rep[:90] + b"NNNN" + rep[3:60], 8),
        ("seed_repeats_before_match", [b"A" * 40 + b"G", b"A" * 33 + b"C"], [b"A" * 20, b"AAAC"],
         b"A" * 40 + b"C" + b"A" * 40 + b"G" + b"A" * 50, 8),
        ("all_T_seed", [b"T" * 40 + b"ACG", b"T" * 36, b"A" * 33 + b"T"], [b"T" * 20, b"TTACG", b"AAAT"],
         b"G" + b"T" * 40 + b"ACG" + b"A" * 40 + b"T", 8),
        ("contig_equals_truth", [rep[:200]], [rep[5:30]], rep[:200], 8),
        ("exact_word_boundaries", [rep[:64], rep[:96], rep[:32]], [rep[:32], rep[32:64], rep[:64], rep[31:64], rep[1:33]], rep, 8),
    ]
