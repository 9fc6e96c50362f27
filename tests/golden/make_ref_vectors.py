"""Generates tests/golden/ref_vectors.npz by running the UNMODIFIED upstream calc_breakscore
(lib/BreakageScorer.cpp:185-353, compiled verbatim by oracle/Makefile into oracle/_ref/) on
small inputs.  Run in the build container, where /root/reference exists:

    make -C oracle && python tests/golden/make_ref_vectors.py

The upstream repository ships no tests or golden vectors for this path;
these vectors are outputs of the reference itself and pin the oracle (tests/test_oracle_golden.py)
and, through it, the CUDA path.  Inputs are stored next to the outputs so that nothing has to be
regenerated on a machine without the reference.

The upstream break-k-mer histogram is internal state; it is pinned through two integer
"checksum tables" (prob[row] = row + 1 and prob[row] = (row + 1)^2 mod 1009 + 1): with them
bp_score = sum_row prob[row] * count[row] is an exact integer in fp64, and for a single placed
read it identifies the incremented row.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from genomeassembler_dev_b200 import synth, tables  # noqa: E402
from oracle import loader as O  # noqa: E402

KMERS = tables.all_kmer_strings()
T = len(KMERS)
REAL = tables.normalised(tables.load_raw_from_csv("/root/reference/data/QueryTable"))
TABLES = {
    "real": REAL,
    "uniform": tables.uniform(T),
    "rowid": np.arange(1, T + 1, dtype=np.float64),
    "rowsq": ((np.arange(1, T + 1, dtype=np.int64) ** 2) % 1009 + 1).astype(np.float64),
}

cases = []


def add(name, path, reads, truth, kmer=8, tabs=("real", "rowid", "rowsq")):
    path = [p if isinstance(p, bytes) else p.encode() for p in path]
    reads = [r if isinstance(r, bytes) else r.encode() for r in reads]
    truth = truth if isinstance(truth, bytes) else truth.encode()
    for tab in tabs:
        res = O.ref_calc_breakscore(path, reads, truth, kmer, KMERS, TABLES[tab], edit_distance=False)
        cases.append(dict(name=f"{name}/{tab}", path=path, reads=reads, truth=truth, kmer=kmer, table=tab, res=res))


# SURVEY.md appendix A.4 known-answer inputs
c = "ACGTTGCAAGGCTTACCGATAGGA"
for pos in range(7):
    add(f"kat1_pos{pos}", [c], [c[pos:pos + 10]] * 3, c)
for pos in (0, 1, 2, 3, 5):
    add(f"kat2_k4_pos{pos}", [c], [c[pos:pos + 10]], c, kmer=4)
add("kat3_leftmost", ["TTTTACGTACGGAAAAACCCCACGTACGGTTTT"], ["ACGTACGG"], "TTTTACGTACGGAAAAACCCCACGTACGGTTTT")
cn = "ACGNTGCAAGGCTTACCGATAGGA"
add("kat4_N", [cn], [cn[5:15]], cn)
add("kat5_startpos4", [c], [c[5:15]], "GGGG" + c + "CC")
add("kat5_startpos_absent", [c], [c[5:15]], "GGGGCC")
add("kat5_nohit", [c], ["TTTTTTTTTT"], "GGGG" + c + "CC")
# contig end clamps the break window (substr past the end), short contigs of 8..11 bases
add("end_clamp", ["ACGTTGCAAG", "ACGTTGCA", "CGTTGCAAGGC"], ["GCAAG", "TTGCA", "CAAGGC", "GTTGC"], c)
# duplicates == weights; read order irrelevant
add("dups", [c, c[3:20]], [c[4:14], c[6:16], c[4:14], c[4:14], c[8:18], c[6:16]], c)

# seeded synthetic segments (velvet-style contigs, mutated contigs, N-gap scaffolds)
for i, (L, r, cov, nc, gaps) in enumerate([(3000, 40, 10, 6, 1), (4000, 100, 8, 5, 0), (2000, 12, 12, 8, 1),
                                          (2500, 150, 10, 4, 1), (1500, 33, 10, 5, 0)]):
    seg = synth.make_segment(100 + i, length=L, read_len=r, coverage=cov, n_contigs=nc, prob8=tables.sub_table(REAL, 8),
                             mut_frac=0.3, n_gap_scaffolds=gaps)
    add(f"synth{i}_L{L}_r{r}", seg.contigs, seg.read_list, seg.truth, tabs=("real", "uniform", "rowid", "rowsq"))
# a repetitive truth: short reads occur many times (leftmost rule)
rng = np.random.default_rng(5)
unit = synth.codes_to_ascii(synth.random_truth_codes(rng, 37)).tobytes()
rep = unit * 30
add("repeats", [rep[5:600], rep[100:400] + b"ACGT" + rep[3:200]], [rep[i:i + 14] for i in range(0, 200, 3)], rep)

out = {}
names = []
for i, cs in enumerate(cases):
    names.append(cs["name"])
    for key in ("path", "reads"):
        ch, off = O.flatten(cs[key])
        out[f"{i}_{key}_chars"] = ch
        out[f"{i}_{key}_off"] = off
    out[f"{i}_truth"] = np.frombuffer(cs["truth"], dtype=np.uint8)
    out[f"{i}_kmer"] = np.int32(cs["kmer"])
    out[f"{i}_table"] = np.array(cs["table"])
    r = cs["res"]
    for key in ("sequence_len", "bp_score", "bp_score_norm_by_break_freqs", "bp_score_norm_by_len", "kmer_breaks",
                "path_prob_dist_startpos"):
        out[f"{i}_{key}"] = r[key]
    out[f"{i}_path_prob_dist"] = np.concatenate(r["path_prob_dist"]) if r["path_prob_dist"] else np.zeros(0)
out["names"] = np.array(names)
dst = os.path.join(HERE, "ref_vectors.npz")
np.savez_compressed(dst, **out)
print(dst, len(cases), "cases", os.path.getsize(dst), "bytes")
