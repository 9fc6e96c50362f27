"""Generates tests/golden/assemble_vectors.npz: outputs of the UNMODIFIED upstream assemble_contigs
(lib/BreakageScorer.cpp:79-174, via oracle/_ref) on small contig sets.  Run where /root/reference exists:
    make -C oracle && python tests/golden/make_assemble_vectors.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from genomeassembler_dev_b200 import synth  # noqa: E402
from oracle import loader as O  # noqa: E402


def overlapping_contigs(seed, length, n, k, dup=False):
    """contigs cut from one random sequence with k-1 .. k+5 bases of overlap between neighbours"""
    rng = np.random.default_rng(seed)
    truth = synth.codes_to_ascii(synth.random_truth_codes(rng, length)).tobytes()
    cuts = np.sort(rng.choice(np.arange(k + 8, length - k - 8), size=n - 1, replace=False))
    bounds = [0] + list(cuts) + [length]
    out = []
    for a, b in zip(bounds[:-1], bounds[1:]):
        ov = int(rng.integers(0, k + 6))
        out.append(truth[max(a - ov, 0):b])
    if dup:
        out.append(out[1])
    order = rng.permutation(len(out))
    return [out[i] for i in order]


CASES = [
    ("six_k13", overlapping_contigs(1, 1200, 6, 13), 13, 1234),
    ("eight_k21", overlapping_contigs(2, 3000, 8, 21), 21, 7),
    ("dup_k11", overlapping_contigs(3, 900, 5, 11, dup=True), 11, 99),
    ("twelve_k15", overlapping_contigs(4, 4000, 12, 15), 15, 1234),
    ("twenty_k31", [c for c in overlapping_contigs(5, 9000, 20, 31) if len(c) >= 30], 31, 1234),
    ("no_overlap", [b"ACGTACGTAAGGCCTT", b"TTGGAACCGGTTAACC", b"GGGGGGGGCCCCCCCC"], 9, 5),
    ("all_equal", [b"ACGTACGTAA"] * 3, 5, 1),
    ("short_contig_throws", [b"ACGTACGTAAGGCCTT", b"ACGT", b"GGGGGGGGCCCCCCCC"], 9, 5),
    ("short_but_all_equal", [b"ACGT", b"ACGT"], 9, 5),
    ("with_empty", [b"ACGTACGTAAGGCCTT", b"", b"GGCCTTACGTTTTTTTTT"], 7, 5),
    ("low_complexity", [b"AAAAAAAAAAAAC", b"CAAAAAAAAAAAA", b"AAAAAAAAAAAAA", b"ACACACACACACA"], 7, 3),
]
out = {"names": np.array([c[0] for c in CASES])}
for i, (name, contigs, k, seed) in enumerate(CASES):
    try:
        res = O.ref_assemble_contigs(contigs, k, seed)
        out[f"{i}_threw"] = np.bool_(False)
    except RuntimeError:  # upstream throws std::out_of_range (substr) when a contig is shorter than dbg_kmer-1
        res = []
        out[f"{i}_threw"] = np.bool_(True)
    for key, strings in (("in", contigs), ("out", res)):
        ch, off = O.flatten(strings)
        out[f"{i}_{key}_chars"], out[f"{i}_{key}_off"] = ch, off
    out[f"{i}_k"], out[f"{i}_seed"] = np.int32(k), np.int32(seed)
    print(name, len(contigs), "->", len(res), [len(x) for x in res[:6]])
np.savez_compressed(os.path.join(HERE, "assemble_vectors.npz"), **out)
