"""Compositional scoring of scaffold sets (upstream assemble_contigs + calc_breakscore on its output,
lib/BreakageScorer.cpp:105-171 and lib/DeNovoAssembler.R:343-355): the host side (compositions out of the scaffold
explosion, validation, texts) without a GPU, and the device algorithm under the CPU emulation.  The GPU runs of the same
cases are in test_gpu_parity.py."""
import os

import numpy as np
import pytest

import scaffold_cases as SC
from test_assemble_contigs import CASES as ASSEMBLE_CASES
from genomeassembler_dev_b200 import breakscore as B


# ---- host side: no GPU -------------------------------------------------------------------------------------

@pytest.mark.parametrize("name,contigs,k,seed,expected,threw", ASSEMBLE_CASES, ids=[c[0] for c in ASSEMBLE_CASES])
def test_assemble_scaffolds_reports_the_parts_of_the_reference_strings(name, contigs, k, seed, expected, threw, product_lib):
    """same strings, same order as the unmodified upstream function; the parts rebuild exactly those strings"""
    if threw:
        with pytest.raises(B.BreakscoreError):
            B.assemble_scaffolds(contigs, k, seed, lib_path=product_lib)
        return
    strings, sset = B.assemble_scaffolds(contigs, k, seed, lib_path=product_lib)
    assert strings == expected
    assert sset.texts() == expected
    assert sset.lengths().tolist() == [len(s) for s in expected]
    # deterministic whatever the thread count
    s1, set1 = B.assemble_scaffolds(contigs, k, seed, n_threads=1, lib_path=product_lib)
    assert s1 == expected
    for f in ("part_start", "part_base", "part_overlap"):
        assert np.array_equal(getattr(set1, f), getattr(sset, f)), f
    # overlaps are what upstream's sweep allows: below dbg_kmer, 0 for a first part
    first = np.zeros(len(sset.part_base), bool)
    first[sset.part_start[:-1]] = True
    assert np.all(sset.part_overlap[first] == 0)
    assert np.all((sset.part_overlap[~first] >= 1) & (sset.part_overlap[~first] < k))


def test_scaffold_set_validation(product_lib):
    base = [b"ACGTACGTAA", b"GTAACCGGTT", b"TTTT"]
    ok = B.ScaffoldSet(base, [0, 2, 3], [0, 1, 2], [0, 4, 0], product_lib)
    assert ok.texts() == [b"ACGTACGTAACCGGTT", b"TTTT"]
    bad = [
        ([0, 2], [0, 1], [0, 3]),      # "TAA" != "GTA": not a suffix/prefix match
        ([0, 2], [0, 1], [2, 4]),      # a first part with an overlap
        ([0, 2], [0, 5], [0, 0]),      # base contig out of range
        ([0, 2], [0, 2], [0, 4]),      # overlap not below the part's length
        ([0, 0, 2], [0, 1], [0, 4]),   # a scaffold without parts
        ([1, 2], [0, 1], [0, 4]),      # part_start[0] != 0
        ([0, 2], [0, 1], [0, -1]),     # negative overlap
    ]
    for ps, pb, po in bad:
        with pytest.raises(B.BreakscoreError):
            B.ScaffoldSet(base, ps, pb, po, product_lib).lengths()
    # an overlap that reaches back over more than one part
    chain = B.ScaffoldSet([b"ACGTAC", b"ACGT", b"CGTAA"], [0, 3], [0, 1, 2], [0, 2, 3], product_lib)
    assert chain.texts() == [b"ACGTAC" + b"GT" + b"AA"]
    # ... but a part may not start before its predecessor (the parts of a scaffold are kept in ascending start order)
    with pytest.raises(B.BreakscoreError):
        B.ScaffoldSet([b"ACGTAC", b"ACG", b"TACGGG"], [0, 3], [0, 1, 2], [0, 2, 4], product_lib).lengths()


# ---- device algorithm under the CPU emulation -------------------------------------------------------------

SETS = [
    dict(seed=61, length=3000, read_len=40, coverage=10, n_base=8, n_scaffolds=30, overlap=9),
    dict(seed=68, length=1500, read_len=8, coverage=6, n_base=6, n_scaffolds=20, overlap=9),    # overlaps longer than the reads: nothing can cross a junction
    dict(seed=62, length=2500, read_len=100, coverage=8, n_base=6, n_scaffolds=20, overlap=0),       # cfg-4 shape: plain concatenations
    dict(seed=63, length=2000, read_len=33, coverage=10, n_base=12, n_scaffolds=25, overlap=15, ragged=True),
    dict(seed=64, length=2600, read_len=150, coverage=10, n_base=10, n_scaffolds=20, overlap=20, mutate=0.5),
]


RANDOM_MODES = ["scored_in_place", "weights", "small_hash", "ks_from_parts", "ks_from_parts_weights", "with_text", "junctions_probed"]
# (CPU suite: every mode on two of the sets, the main modes on the others; the GPU suite runs the full product)
RANDOM_CASES = [(kw, m) for kw in SETS for m in RANDOM_MODES if kw["seed"] in (61, 68) or m in ("scored_in_place", "ks_from_parts")]


@pytest.mark.parametrize("kw,mode", RANDOM_CASES, ids=[f"{m}-seed{k['seed']}" for k, m in RANDOM_CASES])
def test_random_sets_vs_rescan_and_oracle(kw, mode, emul_scorer, emul_lib, oracle, kmers, prob, monkeypatch):
    if mode in ("weights", "ks_from_parts_weights"):
        monkeypatch.setenv("BS_COMPOSE_SCORE", "0")
    if mode == "small_hash":
        monkeypatch.setenv("BS_COMPOSE_HASH_SLOTS", "64")
    if mode == "junctions_probed":  # (with reads of one length the junctions are otherwise placed once per distinct pair)
        monkeypatch.setenv("BS_COMPOSE_JUNCTIONS", "0")
    if mode == "with_text":  # (sets of ACGT only are otherwise kept as packed words alone: k_compose_words)
        monkeypatch.setenv("BS_COMPOSE_TEXT", "1")
    truth, reads, sset = SC.make_set(lib_path=emul_lib, **kw)
    SC.check_scaffolds(emul_scorer, oracle, kmers, prob, truth, reads, sset, flags=SC.mode_flags(mode))


@pytest.mark.parametrize("name,base,chains,reads,truth,kmer", SC.hand_sets(), ids=[h[0] for h in SC.hand_sets()])
@pytest.mark.parametrize("mode", ["scored_in_place", "weights", "global_rows", "small_hash", "ks_from_parts"])
def test_hand_built_sets(name, base, chains, reads, truth, kmer, mode, emul_scorer, emul_lib, oracle, kmers, prob, monkeypatch):
    if mode in ("weights", "ks_from_parts_weights"):
        monkeypatch.setenv("BS_COMPOSE_SCORE", "0")
    if mode == "small_hash":
        monkeypatch.setenv("BS_COMPOSE_HASH_SLOTS", "64")
    if mode == "junctions_probed":  # (with reads of one length the junctions are otherwise placed once per distinct pair)
        monkeypatch.setenv("BS_COMPOSE_JUNCTIONS", "0")
    if mode == "with_text":  # (sets of ACGT only are otherwise kept as packed words alone: k_compose_words)
        monkeypatch.setenv("BS_COMPOSE_TEXT", "1")
    if mode == "global_rows":
        monkeypatch.setenv("BS_COMPOSE_ROWS", "global")
    sset = SC.hand_scaffold_set(base, chains, emul_lib)
    SC.check_scaffolds(emul_scorer, oracle, kmers, prob, truth, reads, sset, kmer=kmer, flags=SC.mode_flags(mode))


def test_second_table_from_the_same_placement(emul_scorer, emul_lib, oracle, kmers, prob):
    from genomeassembler_dev_b200 import tables
    truth, reads, sset = SC.make_set(65, length=2000, read_len=50, coverage=8, n_base=6, n_scaffolds=12, overlap=11, lib_path=emul_lib)
    SC.check_scaffolds(emul_scorer, oracle, kmers, prob, truth, reads, sset, second=tables.uniform(len(prob)))


def test_assembled_set_end_to_end(emul_scorer, emul_lib, oracle, kmers, prob):
    """assemble_scaffolds -> score_scaffolds == assemble_contigs -> calc_breakscore (the upstream sequence of calls)"""
    from genomeassembler_dev_b200 import synth
    seg = synth.make_segment(77, length=2500, read_len=50, coverage=8, n_contigs=7, mut_frac=0.0)
    contigs = [seg.truth[s:s + len(c) + 60] for s, c in zip(seg.contig_truth_start, seg.contigs)]
    strings, sset = B.assemble_scaffolds(contigs, 13, 1234, n_shuffles=300, lib_path=emul_lib)
    assert len(strings) > 3
    got = SC.check_scaffolds(emul_scorer, oracle, kmers, prob, seg.truth, seg.read_list, sset, oracle_sample=range(0, len(strings), 3))
    assert got["sequence"] == strings


def test_refused_flags(emul_scorer, emul_lib, kmers, prob):
    emul_scorer.set_table(kmers, prob)
    sset = SC.hand_scaffold_set([b"ACGTACGTAA", b"GTAACCGGTT"], [[(0, 0), (1, 4)]], emul_lib)
    for f in (B.PLACE_SCAN, B.PLACE_TILE):
        with pytest.raises(B.BreakscoreError):
            emul_scorer.score_scaffolds(sset, [b"ACGT"], b"ACGTACGT", flags=B.DEFAULT_FLAGS | f)
    bad = SC.hand_scaffold_set([b"ACGTACGTAA", b"GTAACCGGTT"], [[(0, 0), (1, 3)]], emul_lib)
    with pytest.raises(B.BreakscoreError):
        emul_scorer.score_scaffolds(bad, [b"ACGT"], b"ACGTACGT")


def test_large_sets_fall_back_to_the_texts(emul_scorer, emul_lib, oracle, kmers, prob, monkeypatch):
    """more (base contig, read) cells than the lists may take: the texts are materialised and scored the ordinary way"""
    monkeypatch.setenv("BS_COMPOSE_MAX_CELLS", "100")
    truth, reads, sset = SC.make_set(66, length=1500, read_len=40, coverage=8, n_base=6, n_scaffolds=10, overlap=9, lib_path=emul_lib)
    before = emul_scorer.launch_count
    SC.check_scaffolds(emul_scorer, oracle, kmers, prob, truth, reads, sset)
    assert emul_scorer.launch_count > before


def long_scaffold_case(lib_path=None):
    """one scaffold of more than 65 536 windows (the rank histogram then keeps 32-bit counters) beside short ones"""
    from genomeassembler_dev_b200 import synth
    rng = np.random.default_rng(70)
    truth = synth.codes_to_ascii(synth.random_truth_codes(rng, 75_000))
    base = [truth[0:26_000].tobytes(), truth[26_000:50_000].tobytes(), truth[50_000:75_000].tobytes(), truth[100:400].tobytes()]
    starts = rng.integers(0, 75_000 - 60, size=300)
    reads = [truth[s:s + 60].tobytes() for s in starts] + [truth[25_990:26_050].tobytes(), truth[49_960:50_020].tobytes()]
    sset = SC.hand_scaffold_set(base, [[(0, 0), (1, 0), (2, 0)], [(3, 0), (1, 0)], [(2, 0)]], lib_path)
    return truth.tobytes(), reads, sset


@pytest.mark.parametrize("mode", ["scored_in_place", "ks_from_parts"])
def test_scaffold_longer_than_65535_windows(mode, emul_scorer, emul_lib, oracle, kmers, prob):
    truth, reads, sset = long_scaffold_case(emul_lib)
    SC.check_scaffolds(emul_scorer, oracle, kmers, prob, truth, reads, sset, flags=SC.mode_flags(mode) & ~B.WANT_POS & ~B.WANT_HIST)


def test_a_scaffolds_record_does_not_depend_on_the_rest_of_the_set(emul_scorer, emul_lib, kmers, prob, monkeypatch):
    SC.check_set_independence(emul_scorer, kmers, prob, monkeypatch, lib_path=emul_lib, seed=71, length=2500, read_len=40, coverage=8,
                              n_base=7, n_scaffolds=11, overlap=9)
