"""Scaffold explosion (upstream assemble_contigs, lib/BreakageScorer.cpp:79-174): the native host
implementation in libbreakscore.so against outputs of the UNMODIFIED upstream function
(tests/golden/assemble_vectors.npz, made by tests/golden/make_assemble_vectors.py) -- same
strings, same order -- and live against oracle/_ref where it is present.  Needs no GPU."""
import os

import numpy as np
import pytest

from conftest import ROOT
from genomeassembler_dev_b200 import breakscore as B


def load_cases():
    z = np.load(os.path.join(ROOT, "tests", "golden", "assemble_vectors.npz"))
    cases = []
    for i, name in enumerate(z["names"]):
        def strings(key):
            ch, off = z[f"{i}_{key}_chars"], z[f"{i}_{key}_off"]
            return [ch[off[j]:off[j + 1]].tobytes() for j in range(len(off) - 1)]
        cases.append((str(name), strings("in"), int(z[f"{i}_k"]), int(z[f"{i}_seed"]), strings("out"), bool(z[f"{i}_threw"])))
    return cases


CASES = load_cases()


@pytest.mark.parametrize("name,contigs,k,seed,expected,threw", CASES, ids=[c[0] for c in CASES])
def test_matches_reference_vectors(name, contigs, k, seed, expected, threw, product_lib):
    if threw:  # upstream: std::out_of_range from substr -> R error
        with pytest.raises(B.BreakscoreError):
            B.assemble_contigs(contigs, k, seed, lib_path=product_lib)
        return
    got = B.assemble_contigs(contigs, k, seed, lib_path=product_lib)
    assert got == expected
    # thread count does not change the result
    assert B.assemble_contigs(contigs, k, seed, n_threads=1, lib_path=product_lib) == expected


def test_live_against_reference_build(oracle, product_lib):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built")
    from genomeassembler_dev_b200 import synth
    seg = synth.make_segment(77, length=2500, read_len=50, coverage=5, n_contigs=7, mut_frac=0.0)
    # neighbouring velvet-style contigs do not overlap: extend each by 12 bases of its successor
    truth = seg.truth
    starts = seg.contig_truth_start
    contigs = [truth[s:s + len(c) + 60] for s, c in zip(starts, seg.contigs)]
    for k, seed in ((13, 1234), (9, 5)):
        assert B.assemble_contigs(contigs, k, seed, lib_path=product_lib) == oracle.ref_assemble_contigs(contigs, k, seed)


def test_feeds_the_scorer_shapes(product_lib):
    """the generator's output is what calc_breakscore takes as `path`: longest first, no duplicates"""
    contigs = CASES[0][1]
    out = B.assemble_contigs(contigs, CASES[0][2], CASES[0][3], lib_path=product_lib)
    assert len(set(out)) == len(out)
    assert all(len(a) >= len(b) for a, b in zip(out, out[1:]))
