"""Parity tests proper: the nvcc-built CUDA library, called through its C-ABI on a B200, against
the oracle (same seeded inputs), the reference vectors, and size-independent properties at
BASELINE.json's full sizes.  Bit-exact for positions / histograms / integer outputs /
path_prob_dist; fp64 scores and KS within 1e-9 relative (conftest.RTOL)."""
import os

import numpy as np
import pytest

import parity_cases as P
from conftest import load_ref_vectors
from genomeassembler_dev_b200 import breakscore as B
from genomeassembler_dev_b200 import synth, tables

pytestmark = pytest.mark.gpu
CASES = load_ref_vectors()


def test_product_library_is_loaded(gpu_scorer, product_lib):
    assert gpu_scorer._lib._name == product_lib
    maps = open("/proc/self/maps").read()
    assert "libbreakscore.so" in maps and "libbreakscore_emul" not in maps.replace("tests/emul", "")


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_reference_vectors(case, gpu_scorer, kmers, table_set):
    P.check_reference_vector(gpu_scorer, case, kmers, table_set)


@pytest.mark.parametrize("params", P.SMALL + P.MEDIUM, ids=[f"L{p[1]}_r{p[2]}" for p in P.SMALL + P.MEDIUM])
@pytest.mark.parametrize("mode", [0, B.PLACE_TILE, B.PLACE_SCAN], ids=["read_index", "tile_index", "scan"])
def test_segments_vs_oracle(params, mode, gpu_scorer, oracle, kmers, prob):
    seg = P.make(*params)
    before = gpu_scorer.launch_count
    P.check_segment(gpu_scorer, oracle, kmers, prob, seg, flags=P.FULL | mode)
    assert gpu_scorer.launch_count > before


@pytest.mark.parametrize("mode", [0, B.PLACE_TILE, B.PLACE_SCAN], ids=["read_index", "tile_index", "scan"])
@pytest.mark.parametrize("name,contigs,reads,truth,kmer", P.edge_inputs(), ids=[e[0] for e in P.edge_inputs()])
def test_edge_inputs(name, contigs, reads, truth, kmer, mode, gpu_scorer, oracle, kmers, prob):
    seg = synth.Segment(truth, None, contigs)
    P.check_segment(gpu_scorer, oracle, kmers, prob, seg, kmer=kmer, reads=reads, flags=P.FULL | mode)


def test_random_pass_keeps_real_truth_table(gpu_scorer, oracle, kmers, prob):
    seg = P.make(41, 20000, 50, 20, 9, 1)
    P.check_segment(gpu_scorer, oracle, kmers, tables.uniform(len(prob)), seg, truth_prob=prob)


def test_long_contig_spans_many_tiles(gpu_scorer, oracle, kmers, prob):
    """contigs longer than one shared-memory tile (scaffolds, cfg-4 shape): leftmost rule across tiles"""
    seg = synth.make_scaffold_set(61, length=30000, read_len=100, coverage=20, n_base=8, n_scaffolds=6, lo=20000, hi=30000)
    P.check_segment(gpu_scorer, oracle, kmers, prob, seg)


def test_batch_equals_oracle_per_segment(gpu_scorer, oracle, kmers, prob):
    gpu_scorer.set_table(kmers, prob)
    b = synth.make_batch(6, seed=70, length=20000, read_len=150, coverage=20, contigs_lo=3, contigs_hi=12, n_gap_scaffolds=1)
    res = gpu_scorer.score_batch(b.read_chars, None, b.read_len, b.contig_chars, b.contig_off, b.truth_chars,
                                 b.truth_off, b.seg_read_start, b.seg_contig_start, flags=B.DEFAULT_FLAGS | B.WANT_HIST)
    for s in range(b.n_segments):
        seg = b.segment(s)
        want = oracle.oracle_calc_breakscore(seg.contigs, seg.read_list, seg.truth, 8, kmers, prob, want_hist=True)
        c0, c1 = int(b.seg_contig_start[s]), int(b.seg_contig_start[s + 1])
        for k in ("sequence_len", "kmer_breaks", "path_prob_dist_startpos", "hist"):
            assert np.array_equal(res[k][c0:c1], want[k]), k
        for k in ("bp_score", "bp_score_norm_by_break_freqs", "bp_score_norm_by_len", "ks_stat_prob_dist", "ks_stat_path_freq"):
            np.testing.assert_allclose(res[k][c0:c1], want[k], rtol=1e-9, atol=1e-12, equal_nan=True, err_msg=k)
        off = res["path_prob_dist_off"]
        for c in range(c0, c1):
            assert np.array_equal(res["path_prob_dist_flat"][off[c]:off[c + 1]], want["path_prob_dist"][c - c0])


# ---- full BASELINE.json sizes: size-independent properties ---------------------------------

@pytest.fixture(scope="module")
def cfg1():
    return synth.make_segment(1234, length=50000, read_len=100, coverage=30, n_contigs=16, mut_frac=0.2, n_gap_scaffolds=1)


def test_cfg1_full_size_vs_oracle(cfg1, gpu_scorer, oracle, kmers, prob):
    P.check_segment(gpu_scorer, oracle, kmers, prob, cfg1)


def test_cfg1_properties(cfg1, gpu_scorer, kmers, prob):
    gpu_scorer.set_table(kmers, prob)
    flags = B.DEFAULT_FLAGS | B.WANT_HIST
    base = gpu_scorer.score(cfg1.contigs, cfg1.reads, cfg1.truth, flags=flags)
    # histogram mass == kmer_breaks; at most three non-octamer rows
    assert np.array_equal(base["hist"].sum(axis=1), base["kmer_breaks"])
    assert (base["hist"][:, :16 + 256 + 4096] > 0).sum(axis=1).max() <= 3
    # read order and contig order are irrelevant; results are bit-identical (fixed reduction order)
    rng = np.random.default_rng(0)
    rp, cp = rng.permutation(len(cfg1.reads)), rng.permutation(len(cfg1.contigs))
    perm = gpu_scorer.score([cfg1.contigs[i] for i in cp], cfg1.reads[rp], cfg1.truth, flags=flags)
    for k in ("kmer_breaks", "bp_score", "bp_score_norm_by_break_freqs", "ks_stat_prob_dist", "ks_stat_path_freq", "hist",
              "path_prob_dist_startpos"):
        assert np.array_equal(perm[k], base[k][cp], equal_nan=True), k
    # duplicates are weights: scoring every read twice doubles counts and bp_score, keeps the normalised score
    dbl = gpu_scorer.score(cfg1.contigs, np.concatenate([cfg1.reads, cfg1.reads]), cfg1.truth, flags=flags)
    assert np.array_equal(dbl["hist"], 2 * base["hist"])
    np.testing.assert_allclose(dbl["bp_score"], 2 * base["bp_score"], rtol=1e-12)
    np.testing.assert_allclose(dbl["bp_score_norm_by_break_freqs"], base["bp_score_norm_by_break_freqs"], rtol=1e-12)
    # the three placement kernels agree bit for bit
    for mode in (B.PLACE_SCAN, B.PLACE_TILE):
        other = gpu_scorer.score(cfg1.contigs, cfg1.reads, cfg1.truth, flags=flags | mode)
        for k in ("kmer_breaks", "bp_score", "hist"):
            assert np.array_equal(other[k], base[k]), k


def test_cfg2_shape_batch_properties(gpu_scorer, oracle, kmers, prob):
    """cfg-2 shape (150 bp, 30x, 50 kb segments) at 24 segments: one batched call == per-segment calls,
    integer-table checksum of the histogram == bp_score, and a sampled segment equals the oracle."""
    b = synth.make_batch(24, seed=1234, length=50000, read_len=150, coverage=30)
    args = (b.read_chars, None, b.read_len, b.contig_chars, b.contig_off, b.truth_chars, b.truth_off,
            b.seg_read_start, b.seg_contig_start)
    T = len(prob)
    rowid = np.arange(1, T + 1, dtype=np.float64)
    gpu_scorer.set_table(kmers, rowid)
    chk = gpu_scorer.score_batch(*args, flags=B.WANT_HIST)
    assert np.array_equal(chk["bp_score"], (chk["hist"][:, :T].astype(np.float64) * rowid).sum(axis=1))
    gpu_scorer.set_table(kmers, prob)
    res = gpu_scorer.score_batch(*args, flags=B.DEFAULT_FLAGS)
    assert np.array_equal(res["kmer_breaks"], chk["kmer_breaks"])
    for s in (0, 11, 23):
        seg = b.segment(s)
        want = oracle.oracle_calc_breakscore(seg.contigs, seg.read_list, seg.truth, 8, kmers, prob)
        c0, c1 = int(b.seg_contig_start[s]), int(b.seg_contig_start[s + 1])
        assert np.array_equal(res["kmer_breaks"][c0:c1], want["kmer_breaks"])
        assert np.array_equal(res["path_prob_dist_startpos"][c0:c1], want["path_prob_dist_startpos"])
        np.testing.assert_allclose(res["bp_score"][c0:c1], want["bp_score"], rtol=1e-9)
        np.testing.assert_allclose(res["ks_stat_prob_dist"][c0:c1], want["ks_stat_prob_dist"], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(res["ks_stat_path_freq"][c0:c1], want["ks_stat_path_freq"], rtol=1e-9, atol=1e-12, equal_nan=True)


def test_argument_errors(gpu_scorer, kmers, prob):
    with pytest.raises(B.BreakscoreError):
        gpu_scorer.set_table(["ACGX"], [0.5])
    gpu_scorer.set_table(kmers, prob)
    with pytest.raises(B.BreakscoreError):
        gpu_scorer.score([b"ACGT"], [b"AC"], b"ACGT", kmer=0)


def test_chunked_pipeline_equals_one_chunk(product_lib, gpu_scorer, kmers, prob, monkeypatch):
    """the H2D / compute / D2H pipeline over many chunks == one chunk, bit for bit"""
    b = synth.make_batch(40, seed=90, length=20000, read_len=100, coverage=20, contigs_lo=2, contigs_hi=10)
    args = (b.read_chars, None, b.read_len, b.contig_chars, b.contig_off, b.truth_chars, b.truth_off,
            b.seg_read_start, b.seg_contig_start)
    flags = B.DEFAULT_FLAGS | B.WANT_HIST
    gpu_scorer.set_table(kmers, prob)
    one = gpu_scorer.score_batch(*args, flags=flags)
    monkeypatch.setenv("BS_CHUNK_KB", "1500")
    with B.BreakageScorer(0, product_lib) as sc:
        sc.set_table(kmers, prob)
        many = sc.score_batch(*args, flags=flags)
        again = sc.score_batch(*args, flags=flags)
    for k in one:
        assert np.array_equal(one[k], many[k], equal_nan=True), k
        assert np.array_equal(one[k], again[k], equal_nan=True), k


def test_poll_callback_interrupts_between_chunks(product_lib, kmers, prob, monkeypatch):
    """bs_ctx_set_poll on the GPU pipeline: polled between chunks, a true return ends the call with
    BS_ERR_INTERRUPTED after the chunks in flight have finished; the context stays usable."""
    b = synth.make_batch(40, seed=91, length=20000, read_len=100, coverage=20, contigs_lo=2, contigs_hi=10)
    args = (b.read_chars, None, b.read_len, b.contig_chars, b.contig_off, b.truth_chars, b.truth_off,
            b.seg_read_start, b.seg_contig_start)
    monkeypatch.setenv("BS_CHUNK_KB", "1500")
    with B.BreakageScorer(0, product_lib) as sc:
        sc.set_table(kmers, prob)
        want = sc.score_batch(*args)
        calls = []
        sc.set_poll(lambda: calls.append(1) or len(calls) >= 2)
        with pytest.raises(B.BreakscoreError) as ei:
            sc.score_batch(*args)
        assert ei.value.code == B.ERR_INTERRUPTED and len(calls) == 2
        sc.set_poll(None)
        again = sc.score_batch(*args)
    for k in want:
        assert np.array_equal(want[k], again[k], equal_nan=True), k


@pytest.mark.parametrize("p2p_mb", [None, "0"], ids=["reads_over_pcie", "reads_by_peer_copies"])
def test_one_segment_over_several_contexts(p2p_mb, product_lib, gpu_scorer, kmers, prob, monkeypatch):
    """bs_score_multi on the GPU: three contexts (here on the same device; on an 8-GPU box one per GPU), a host
    thread each, contigs dealt out longest first, reads replicated == one call on one context, byte for byte.
    BS_MULTI_P2P_MB=0: the reads go to context 0 once and reach the others by peer copies (what a large read set does)."""
    if p2p_mb is not None:
        monkeypatch.setenv("BS_MULTI_P2P_MB", p2p_mb)
    seg = P.make(77, 20000, 64, 20, 25, 1, mut=0.3)
    contigs = list(seg.contigs) + [seg.contigs[0][:5], b"", seg.contigs[-1]]
    flags = B.DEFAULT_FLAGS | B.WANT_HIST | B.WANT_POS | B.WANT_LEV
    gpu_scorer.set_table(kmers, prob)
    want = gpu_scorer.score(contigs, seg.reads, seg.truth, flags=flags)
    others = [B.BreakageScorer(0, product_lib) for _ in range(2)]
    try:
        for o in others:
            o.set_table(kmers, prob)
        got = gpu_scorer.score(contigs, seg.reads, seg.truth, flags=flags, group=others)
    finally:
        for o in others:
            o.close()
    for k in want:
        if k == "path_prob_dist":
            assert all(np.array_equal(x, y) for x, y in zip(want[k], got[k])), k
        elif k != "sequence":
            assert np.array_equal(want[k], got[k], equal_nan=True), k


# ---- cfg-4 / cfg-5 shapes ---------------------------------------------------------------------

def test_cfg4_scaffold_set(gpu_scorer, oracle, kmers, prob):
    """cfg-4 shape: thousands of 10-50 kb scaffolds (permuted concatenations of the same base contigs)
    of ONE segment.  A sample of scaffolds against the oracle; all of them through size-independent
    properties: histogram checksum under the integer table, equal scaffolds score equally, and
    contig-sharding over 1/2/4/8 ranks (reads replicated) changes nothing, bit for bit."""
    from genomeassembler_dev_b200 import sharding
    seg = synth.make_scaffold_set(404, length=50000, read_len=150, coverage=30, n_base=16, n_scaffolds=1500)
    sample = list(range(0, 1500, 75))
    sub = synth.Segment(seg.truth, seg.reads, [seg.contigs[i] for i in sample])
    P.check_segment(gpu_scorer, oracle, kmers, prob, sub, flags=B.DEFAULT_FLAGS | B.WANT_HIST)
    T = len(prob)
    rowid = np.arange(1, T + 1, dtype=np.float64)
    gpu_scorer.set_table(kmers, rowid)
    chk = gpu_scorer.score(seg.contigs, seg.reads, seg.truth, flags=B.WANT_HIST)
    assert np.array_equal(chk["bp_score"], (chk["hist"][:, :T].astype(np.float64) * rowid).sum(axis=1))
    assert np.array_equal(chk["hist"].sum(axis=1), chk["kmer_breaks"])
    gpu_scorer.set_table(kmers, prob)
    flags = B.WANT_KS | B.WANT_STARTPOS
    whole = gpu_scorer.score(seg.contigs, seg.reads, seg.truth, flags=flags)
    assert np.array_equal(whole["kmer_breaks"], chk["kmer_breaks"])
    first = {}
    for i, c in enumerate(seg.contigs):
        j = first.setdefault(c, i)
        if j != i:
            for k in ("bp_score", "kmer_breaks", "ks_stat_prob_dist", "ks_stat_path_freq", "path_prob_dist_startpos"):
                assert whole[k][i] == whole[k][j] or (np.isnan(whole[k][i]) and np.isnan(whole[k][j])), k
    lens = [len(c) for c in seg.contigs]
    for world in (2, 4, 8):
        parts = sharding.shard_contigs_lpt(lens, world)
        for part in parts:
            res = gpu_scorer.score([seg.contigs[i] for i in part], seg.reads, seg.truth, flags=flags)
            for k in ("bp_score", "bp_score_norm_by_break_freqs", "kmer_breaks", "ks_stat_prob_dist", "ks_stat_path_freq",
                      "path_prob_dist_startpos"):
                assert np.array_equal(res[k], whole[k][part], equal_nan=True), (world, k)


def test_cfg4_full_size_sampled_oracle(gpu_scorer, kmers, prob):
    """BASELINE.json configs[3] at its size: 10 000 scaffolds in ONE call.  208 of them -- the longest, the shortest,
    every member of a few groups of identical scaffolds, the rest at random -- against the CPU oracle with all reads
    (contigs are independent: lib/BreakageScorer.cpp:231-304), straight out of the 10 000-scaffold call's own output."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import loader as O
    seg = synth.make_scaffold_set(1400, n_scaffolds=10000)
    gpu_scorer.set_table(kmers, prob)
    res = gpu_scorer.score(seg.contigs, seg.reads, seg.truth, flags=B.WANT_KS | B.WANT_STARTPOS)
    lens = np.array([len(c) for c in seg.contigs])
    groups = {}
    for i, c in enumerate(seg.contigs):
        groups.setdefault(c, []).append(i)
    dups = [g for g in groups.values() if len(g) > 1][:8]
    chosen = {int(np.argmax(lens)), int(np.argmin(lens))}
    for g in dups:
        chosen.update(g[:4])
    rng = np.random.default_rng(44)
    for i in rng.permutation(len(seg.contigs)):
        if len(chosen) >= 208:
            break
        chosen.add(int(i))
    chosen = sorted(chosen)
    reads = seg.read_list

    def check(idx):
        want = O.oracle_calc_breakscore([seg.contigs[i] for i in idx], reads, seg.truth, 8, kmers, prob, want_prob_dist=False)
        for k in ("sequence_len", "kmer_breaks", "path_prob_dist_startpos"):
            assert np.array_equal(res[k][idx], want[k]), k
        for k in ("bp_score", "bp_score_norm_by_break_freqs", "bp_score_norm_by_len"):
            np.testing.assert_allclose(res[k][idx], want[k], rtol=1e-9, atol=0, err_msg=k)
        for k in ("ks_stat_prob_dist", "ks_stat_path_freq"):
            np.testing.assert_allclose(res[k][idx], want[k], rtol=1e-9, atol=1e-12, equal_nan=True, err_msg=k)
        return len(idx)

    with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as ex:
        assert sum(ex.map(check, [chosen[i:i + 4] for i in range(0, len(chosen), 4)])) >= 200
    for g in dups:  # identical scaffolds score identically
        for k in ("bp_score", "kmer_breaks", "ks_stat_prob_dist"):
            assert len({res[k][i].tobytes() for i in g}) == 1, k


def test_cfg5_shape_scaled(gpu_scorer, oracle, kmers, prob):
    """cfg-5 shape scaled to test size: one 1 Mb truth, 1e5 uniform-start 150 bp reads, 1000 contigs of
    about 1 kb.  A sample of contigs against the oracle (contigs are independent), the rest through the
    histogram checksum."""
    rng = np.random.default_rng(505)
    L, N, Cn, r = 1_000_000, 100_000, 1000, 150
    truth = synth.codes_to_ascii(synth.random_truth_codes(rng, L))
    starts = rng.integers(0, L - r, size=N)
    reads = truth[starts[:, None] + np.arange(r)[None, :]]
    cstart = np.sort(rng.integers(0, L - 3000, size=Cn))
    clen = rng.integers(200, 2000, size=Cn)
    contigs = [truth[a:a + b].tobytes() for a, b in zip(cstart, clen)]
    seg = synth.Segment(truth.tobytes(), reads, contigs)
    sample = list(range(0, Cn, 50))
    sub = synth.Segment(seg.truth, seg.reads, [contigs[i] for i in sample])
    got, want = P.check_segment(gpu_scorer, oracle, kmers, prob, sub, flags=B.DEFAULT_FLAGS | B.WANT_HIST)
    assert np.array_equal(got["path_prob_dist_startpos"], cstart[sample].astype(np.int32) * (got["kmer_breaks"] > 0))
    T = len(prob)
    rowid = np.arange(1, T + 1, dtype=np.float64)
    gpu_scorer.set_table(kmers, rowid)
    chk = gpu_scorer.score(contigs, reads, seg.truth, flags=B.WANT_HIST | B.WANT_STARTPOS)
    assert np.array_equal(chk["bp_score"], (chk["hist"][:, :T].astype(np.float64) * rowid).sum(axis=1))
    # a read starting inside a contig and ending inside it is placed there exactly once
    expect = np.array([np.count_nonzero((starts >= a) & (starts + r <= a + b)) for a, b in zip(cstart, clen)])
    assert np.all(chk["kmer_breaks"] >= expect)


def test_startpos_of_a_thousand_contigs_in_one_segment(gpu_scorer, kmers, prob):
    """cfg-5 shape: every contig of a 1001-contig segment is an exact substring of the 1 Mb truth at a known
    offset (the seed tables of the contig-in-truth search hold 1001 seeds; 489 blocks scan the truth).  Repeated:
    an earlier kernel lost about 1 % of these offsets at random."""
    rng = np.random.default_rng(506)
    L, N, Cn, r = 1_000_000, 100_000, 1000, 150
    truth = synth.codes_to_ascii(synth.random_truth_codes(rng, L))
    starts = rng.integers(0, L - r, size=N)
    reads = truth[starts[:, None] + np.arange(r)[None, :]]
    cstart = np.sort(rng.integers(0, L - 3000, size=Cn))
    contigs = [truth[a:a + b].tobytes() for a, b in zip(cstart, rng.integers(200, 2000, size=Cn))]
    contigs.append(truth[5000:65000].tobytes())
    contigs.append(truth[5000:6000].tobytes() + b"A" + truth[6001:7000].tobytes())   # same seed as the one before, not a substring
    contigs.append(truth[700000:700100].tobytes())                                   # inside other contigs or not: found at its own offset
    want = np.concatenate([cstart, [5000, -1, 700000]]).astype(np.int32)
    if truth[6000] == ord("A"):
        want[-2] = 5000
    gpu_scorer.set_table(kmers, prob)
    for it in range(6):
        res = gpu_scorer.score(contigs, reads, truth.tobytes(), flags=B.WANT_STARTPOS if it % 2 else B.DEFAULT_FLAGS)
        assert np.array_equal(res["path_prob_dist_startpos"], want * (res["kmer_breaks"] > 0)), it
        assert np.count_nonzero(res["kmer_breaks"]) >= 1000


def test_startpos_many_contigs(gpu_scorer, oracle, kmers, prob):
    from test_emul_device_algorithm import check_startpos_many_contigs
    for _ in range(3):
        check_startpos_many_contigs(gpu_scorer, oracle, kmers, prob, n_contigs=6000, L=200000)


def test_startpos_table_geometries(gpu_scorer, oracle, kmers, prob, monkeypatch):
    from test_emul_device_algorithm import check_startpos_geometries
    check_startpos_geometries(gpu_scorer, oracle, kmers, prob, monkeypatch, n_contigs=6000, L=200000)


# ---- infix edit distance (lev_dist_vs_true) ---------------------------------------------------

@pytest.mark.parametrize("params", P.SMALL + P.MEDIUM[:2], ids=[f"L{p[1]}_r{p[2]}" for p in P.SMALL + P.MEDIUM[:2]])
def test_edit_distance_vs_oracle(params, gpu_scorer, oracle, kmers, prob):
    seg = P.make(*params, mut=0.5)
    P.check_segment(gpu_scorer, oracle, kmers, prob, seg, flags=P.FULL | B.WANT_LEV)


@pytest.mark.parametrize("name,contigs,reads,truth,kmer", P.edge_inputs(), ids=[e[0] for e in P.edge_inputs()])
def test_edit_distance_edge_inputs(name, contigs, reads, truth, kmer, gpu_scorer, oracle, kmers, prob):
    P.check_segment(gpu_scorer, oracle, kmers, prob, synth.Segment(truth, None, contigs), kmer=kmer, reads=reads,
                    flags=B.DEFAULT_FLAGS | B.WANT_LEV)


def test_edit_distance_indels_and_long_contigs(gpu_scorer, oracle, kmers, prob):
    rng = np.random.default_rng(9)
    truth = synth.codes_to_ascii(synth.random_truth_codes(rng, 20000)).tobytes()
    a = bytearray(truth[100:9700]); a[4300] = ord("A") if a[4300] != ord("A") else ord("C")
    b = truth[300:1500] + truth[1510:8900]
    c = truth[2000:3000] + b"ACGTTGCA" + truth[3000:14700]
    d = synth.codes_to_ascii(synth.random_truth_codes(rng, 3000)).tobytes()
    e = truth[14000:20000] + b"GGGGGGGGGG"
    seg = synth.Segment(truth, None, [bytes(a), b, c, d, e, truth[50:12500]])
    reads = [truth[i:i + 100] for i in range(0, 19000, 97)]
    got, want = P.check_segment(gpu_scorer, oracle, kmers, prob, seg, reads=reads, flags=B.DEFAULT_FLAGS | B.WANT_LEV)
    assert got["lev_dist_vs_true"].tolist()[:3] == [1, 10, 8] and got["lev_dist_vs_true"][5] == 0


def test_cfg2_full_study_properties(gpu_scorer, kmers, prob):
    """BASELINE.json configs[1] at full size (1000 segments, ~1e7 reads, ~3e4 contigs): the integer
    checksum table gives integer scores and the same break counts as the real table; scoring the study
    in one call, in two halves or in many pipeline chunks is bit-identical; lengths and offsets agree
    with the inputs."""
    b = synth.make_batch(1000, seed=1234, length=50000, read_len=150, coverage=30)
    args = (b.read_chars, None, b.read_len, b.contig_chars, b.contig_off, b.truth_chars, b.truth_off,
            b.seg_read_start, b.seg_contig_start)
    T = len(prob)
    gpu_scorer.set_table(kmers, np.arange(1, T + 1, dtype=np.float64))
    chk = gpu_scorer.score_batch(*args, flags=B.WANT_STARTPOS)
    assert np.array_equal(chk["bp_score"], np.round(chk["bp_score"])) and chk["bp_score"].max() < 2.0 ** 53
    assert np.array_equal(chk["sequence_len"], np.diff(b.contig_off))
    gpu_scorer.set_table(kmers, prob)
    whole = gpu_scorer.score_batch(*args, flags=B.WANT_KS | B.WANT_STARTPOS)
    assert np.array_equal(whole["kmer_breaks"], chk["kmer_breaks"])
    assert np.array_equal(whole["path_prob_dist_startpos"], chk["path_prob_dist_startpos"])
    assert whole["kmer_breaks"].sum() > 0.8 * b.n_reads   # almost every read lands in exactly one contig of its segment
    # exact substrings report their cut position, mutated contigs -1 (make_segment's bookkeeping)
    assert (whole["path_prob_dist_startpos"] >= -1).all()
    from genomeassembler_dev_b200 import sharding
    keys = ("bp_score", "bp_score_norm_by_break_freqs", "kmer_breaks", "ks_stat_prob_dist", "ks_stat_path_freq", "path_prob_dist_startpos")
    for s0, s1 in sharding.shard_segments(b.seg_read_start, b.seg_contig_start, b.contig_off, 2):
        part = gpu_scorer.score_batch(*sharding.slice_batch(b, s0, s1), flags=B.WANT_KS | B.WANT_STARTPOS)
        c0, c1 = int(b.seg_contig_start[s0]), int(b.seg_contig_start[s1])
        for k in keys:
            assert np.array_equal(part[k], whole[k][c0:c1], equal_nan=True), k
    # 32 randomly chosen segments of THIS 1000-segment call's own output (default outputs) against the unmodified
    # reference build where it exists (oracle/_ref travels with the repo), else the C restatement: integer columns and
    # path_prob_dist bit-exact, scores 1e-9; the KS columns against the restatement (the reference has none)
    from concurrent.futures import ThreadPoolExecutor
    from oracle import loader as O
    full = gpu_scorer.score_batch(*args, flags=B.DEFAULT_FLAGS)
    for k in keys:
        assert np.array_equal(full[k], whole[k], equal_nan=True), k
    pick = sorted(np.random.default_rng(20261019).choice(b.n_segments, size=32, replace=False).tolist())

    def check(s):
        seg = b.segment(s)
        c0, c1 = int(b.seg_contig_start[s]), int(b.seg_contig_start[s + 1])
        port = O.oracle_calc_breakscore(seg.contigs, seg.read_list, seg.truth, 8, kmers, prob)
        want = O.ref_calc_breakscore(seg.contigs, seg.read_list, seg.truth, 8, kmers, prob) if O.have_ref() else port
        for k in ("sequence_len", "kmer_breaks", "path_prob_dist_startpos"):
            assert np.array_equal(full[k][c0:c1], want[k]), (s, k)
        for k in ("bp_score", "bp_score_norm_by_break_freqs", "bp_score_norm_by_len"):
            np.testing.assert_allclose(full[k][c0:c1], want[k], rtol=1e-9, atol=0, err_msg=f"{s} {k}")
        pd0, pd1 = int(full["path_prob_dist_off"][c0]), int(full["path_prob_dist_off"][c1])
        assert np.array_equal(full["path_prob_dist_flat"][pd0:pd1], np.concatenate(want["path_prob_dist"])), (s, "path_prob_dist")
        for k in ("ks_stat_prob_dist", "ks_stat_path_freq"):
            np.testing.assert_allclose(full[k][c0:c1], port[k], rtol=1e-9, atol=1e-12, equal_nan=True, err_msg=f"{s} {k}")
        return c1 - c0

    with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as ex:  # (the ctypes calls release the GIL)
        assert sum(ex.map(check, pick)) > 300


@pytest.mark.parametrize("mode", ["default", "tile", "lev", "second_table", "ragged", "big_startpos"])
def test_hundred_repeats_are_bit_identical(mode, gpu_scorer, oracle, kmers, prob, monkeypatch):
    """Every kernel that uses warp-level primitives or inter-block atomics (k_place ballot counts, the KS prefix sums and
    max reductions, block_sum_fixed, k_startpos_verify, k_lev_bound / k_lev_infix, the packing and index kernels),
    100 calls on the same inputs: every output of every call equals the first call's bit for bit, and the first call's
    equals the oracle's.  (An earlier contig-in-truth kernel lost offsets at random on the GPU only: DESIGN.md section 5.)"""
    gpu_scorer.set_table(kmers, prob)
    flags = B.DEFAULT_FLAGS | B.WANT_HIST
    if mode == "big_startpos":  # 1500 contigs in one segment: one-table-per-segment geometry, candidate queue, warp verification
        rng = np.random.default_rng(9)
        L = 300_000
        truth = synth.codes_to_ascii(synth.random_truth_codes(rng, L))
        starts = rng.integers(0, L - 100, size=60_000)
        reads = truth[starts[:, None] + np.arange(100)[None, :]]
        cs = np.sort(rng.integers(0, L - 41000, size=1500))
        cl = rng.integers(100, 1500, size=1500)
        cl[::300] = 40000  # a few long contigs: many verification steps per candidate
        contigs = [truth[a:a + b].tobytes() for a, b in zip(cs, cl)]
        seg = synth.Segment(truth.tobytes(), reads, contigs)
        flags = B.WANT_STARTPOS | B.WANT_KS
        call = lambda: gpu_scorer.score(seg.contigs, seg.reads, seg.truth, flags=flags)  # noqa: E731
        first = call()
        assert np.array_equal(first["path_prob_dist_startpos"], cs.astype(np.int32) * (first["kmer_breaks"] > 0))
    else:
        b = synth.make_batch(12, seed=777, length=20000, read_len=100, coverage=20, contigs_lo=3, contigs_hi=30, n_gap_scaffolds=1)
        args = [b.read_chars, None, b.read_len, b.contig_chars, b.contig_off, b.truth_chars, b.truth_off, b.seg_read_start, b.seg_contig_start]
        if mode == "tile":
            flags |= B.PLACE_TILE
        elif mode == "lev":
            flags |= B.WANT_LEV
        elif mode == "second_table":
            gpu_scorer.set_second_table(tables.uniform(len(prob)))
            flags |= B.WANT_SECOND_TABLE
        elif mode == "ragged":  # offsets given and not uniform: k_pack_reads
            n = b.n_reads
            lens = np.full(n, b.read_len, np.int64)
            lens[::7] = 60
            off = np.zeros(n + 1, np.int64)
            np.cumsum(lens, out=off[1:])
            rows = b.read_chars.reshape(n, b.read_len)
            args[0] = np.concatenate([rows[i, :lens[i]] for i in range(n)])
            args[1], args[2] = off, 0
        call = lambda: gpu_scorer.score_batch(*args, flags=flags)  # noqa: E731
        first = call()
        s = 5
        seg = b.segment(s)
        rl = seg.read_list if mode != "ragged" else [args[0][args[1][i]:args[1][i + 1]].tobytes() for i in range(int(b.seg_read_start[s]), int(b.seg_read_start[s + 1]))]
        want = oracle.oracle_calc_breakscore(seg.contigs, rl, seg.truth, 8, kmers, prob, want_hist=True, want_lev=(mode == "lev"))
        c0, c1 = int(b.seg_contig_start[s]), int(b.seg_contig_start[s + 1])
        for k in ("kmer_breaks", "path_prob_dist_startpos", "hist") + (("lev_dist_vs_true",) if mode == "lev" else ()):
            assert np.array_equal(first[k][c0:c1], want[k]), k
        np.testing.assert_allclose(first["ks_stat_prob_dist"][c0:c1], want["ks_stat_prob_dist"], rtol=1e-9, atol=1e-12)
    for it in range(int(os.environ.get("BS_TEST_REPS", "100"))):
        again = call()
        for k in first:
            if k != "sequence":
                assert np.array_equal(first[k], again[k], equal_nan=True), (it, k)
    if mode == "second_table":
        gpu_scorer.set_second_table(None)


def test_fused_scoring(gpu_scorer, oracle, kmers, prob, monkeypatch):
    from test_emul_device_algorithm import check_fused_scoring
    check_fused_scoring(gpu_scorer, oracle, kmers, prob, monkeypatch, long_len=16400)
    check_fused_scoring(gpu_scorer, oracle, kmers, prob, monkeypatch, long_len=40000)


def test_pack_variants(gpu_scorer, kmers, prob, monkeypatch):
    """bulk-copy staged and register staged packing out of device buffers at every 16-byte phase (exact-size
    torch tensors: the bulk copy must not touch a byte outside them)"""
    import torch
    from test_emul_device_algorithm import check_pack_variants
    check_pack_variants(gpu_scorer, kmers, prob, monkeypatch, [1, 12, 31, 33, 100, 150, 151, 1000, 3000],
                        to_dev=lambda a: torch.from_numpy(a).cuda(), n_reads=(1, 3, 257, 1111, 20011))


def test_async_device_resident_calls_reuse_workspaces(product_lib, kmers, prob, monkeypatch):
    """BS_DEVICE_CHARS | BS_DEVICE_RESULT calls return before their kernels finish; back-to-back calls
    over many small chunks must not rewrite a workspace that a running chunk still reads."""
    import torch
    b1 = synth.make_batch(30, seed=300, length=20000, read_len=100, coverage=20, contigs_lo=2, contigs_hi=10)
    b2 = synth.make_batch(30, seed=900, length=20000, read_len=100, coverage=20, contigs_lo=2, contigs_hi=10)
    monkeypatch.setenv("BS_CHUNK_KB", "2000")
    with B.BreakageScorer(0, product_lib) as sc:
        sc.set_table(kmers, prob)
        stream = torch.cuda.Stream()
        sc.set_stream(stream.cuda_stream)
        outs = []
        with torch.cuda.stream(stream):
            for rep in range(3):
                for b in (b1, b2):
                    dev = [torch.from_numpy(x).cuda() for x in (b.read_chars, b.contig_chars, b.truth_chars)]
                    Cn = b.n_contigs
                    i32 = torch.zeros(4, Cn, dtype=torch.int32, device="cuda")
                    f64 = torch.zeros(5, Cn, dtype=torch.float64, device="cuda")
                    bt = B._Batch(b.n_segments, b.n_reads, Cn, dev[0].data_ptr(), None, b.read_len, dev[1].data_ptr(),
                                  b.contig_off.ctypes.data, dev[2].data_ptr(), b.truth_off.ctypes.data,
                                  b.seg_read_start.ctypes.data, b.seg_contig_start.ctypes.data)
                    r = B._Result()
                    r.sequence_len, r.kmer_breaks, r.path_prob_dist_startpos, r.lev_dist_vs_true = [i32[i].data_ptr() for i in range(4)]
                    (r.bp_score, r.bp_score_norm_by_break_freqs, r.bp_score_norm_by_len, r.ks_stat_prob_dist,
                     r.ks_stat_path_freq) = [f64[i].data_ptr() for i in range(5)]
                    sc.score_batch_raw(bt, r, 8, B.WANT_KS | B.WANT_STARTPOS | B.DEVICE_CHARS | B.DEVICE_RESULT)
                    outs.append((b, dev, i32, f64))
        sc.synchronize()
        torch.cuda.synchronize()
        for b, dev, i32, f64 in outs:
            ref = sc.score_batch(b.read_chars, None, b.read_len, b.contig_chars, b.contig_off, b.truth_chars, b.truth_off,
                                 b.seg_read_start, b.seg_contig_start, flags=B.WANT_KS | B.WANT_STARTPOS)
            assert np.array_equal(i32[1].cpu().numpy(), ref["kmer_breaks"])
            assert np.array_equal(i32[2].cpu().numpy(), ref["path_prob_dist_startpos"])
            assert np.array_equal(f64[0].cpu().numpy(), ref["bp_score"])
            assert np.array_equal(f64[3].cpu().numpy(), ref["ks_stat_prob_dist"], equal_nan=True)


def test_second_table_in_one_call(gpu_scorer, kmers, prob):
    from test_emul_device_algorithm import check_second_table
    check_second_table(gpu_scorer, kmers, prob, P.make(44, 50000, 150, 30, 20, 1))


def test_irregular_batch(product_lib, gpu_scorer, oracle, kmers, prob, monkeypatch):
    from test_emul_device_algorithm import check_irregular_batch
    one = check_irregular_batch(gpu_scorer, oracle, kmers, prob)
    monkeypatch.setenv("BS_CHUNK_KB", "3")
    with B.BreakageScorer(0, product_lib) as sc:
        many = check_irregular_batch(sc, oracle, kmers, prob)
    for k in one:
        assert np.array_equal(one[k], many[k], equal_nan=True), k


def test_uniform_read_lengths(gpu_scorer, oracle, kmers, prob):
    from test_emul_device_algorithm import UNIFORM_LENGTHS, check_uniform_read_lengths
    check_uniform_read_lengths(gpu_scorer, oracle, kmers, prob, UNIFORM_LENGTHS + [250, 1000])


def test_spectrum_variants(gpu_scorer, product_lib, kmers, prob):
    from test_emul_device_algorithm import check_spectrum_variants
    check_spectrum_variants(gpu_scorer, product_lib, kmers, prob)


# ---- hashed placement scratch (segments with too many reads for a dense row per block, cfg-5) ----

@pytest.mark.parametrize("params", P.SMALL[:3] + [P.MEDIUM[0], P.MEDIUM[1], P.MEDIUM[3]],
                         ids=[f"L{p[1]}_r{p[2]}" for p in P.SMALL[:3] + [P.MEDIUM[0], P.MEDIUM[1], P.MEDIUM[3]]])
def test_hashed_scratch_vs_oracle(params, gpu_scorer, oracle, kmers, prob, monkeypatch):
    from test_emul_device_algorithm import check_hashed_scratch
    check_hashed_scratch(gpu_scorer, oracle, kmers, prob, P.make(*params), monkeypatch, block_threads=256)


@pytest.mark.parametrize("name,contigs,reads,truth,kmer", P.edge_inputs(), ids=[e[0] for e in P.edge_inputs()])
def test_hashed_scratch_edge_inputs(name, contigs, reads, truth, kmer, gpu_scorer, oracle, kmers, prob, monkeypatch):
    monkeypatch.setenv("BS_PLACE_SCRATCH_MB", "0")
    monkeypatch.setenv("BS_PLACE_HASH_CAP", "1")
    P.check_segment(gpu_scorer, oracle, kmers, prob, synth.Segment(truth, None, contigs), kmer=kmer, reads=reads)


def test_hashed_scratch_cfg5_shape(gpu_scorer, kmers, prob, monkeypatch):
    """cfg-5 shape (one 1 Mb truth, 1e5 reads, 1000 contigs) and a batch of segments: the hashed scratch gives
    the arrays of the dense scratch, bit for bit"""
    rng = np.random.default_rng(506)
    L, N, Cn, r = 1_000_000, 100_000, 1000, 150
    truth = synth.codes_to_ascii(synth.random_truth_codes(rng, L))
    starts = rng.integers(0, L - r, size=N)
    reads = truth[starts[:, None] + np.arange(r)[None, :]]
    cstart = np.sort(rng.integers(0, L - 3000, size=Cn))
    contigs = [truth[a:a + b].tobytes() for a, b in zip(cstart, rng.integers(200, 2000, size=Cn))]
    contigs.append(truth[5000:65000].tobytes())   # one long scaffold: thousands of reads in one table
    b = synth.make_batch(6, seed=17, length=20000, read_len=100, coverage=20, contigs_lo=3, contigs_hi=9)
    bargs = (b.read_chars, None, b.read_len, b.contig_chars, b.contig_off, b.truth_chars, b.truth_off,
             b.seg_read_start, b.seg_contig_start)
    flags = B.DEFAULT_FLAGS | B.WANT_HIST | B.WANT_POS
    gpu_scorer.set_table(kmers, prob)
    dense = gpu_scorer.score(contigs, reads, truth.tobytes(), flags=flags)
    dense_b = gpu_scorer.score_batch(*bargs, flags=flags)
    monkeypatch.setenv("BS_PLACE_SCRATCH_MB", "0")
    for cap in (None, "1"):
        if cap:
            monkeypatch.setenv("BS_PLACE_HASH_CAP", cap)
        hashed = gpu_scorer.score(contigs, reads, truth.tobytes(), flags=flags)
        hashed_b = gpu_scorer.score_batch(*bargs, flags=flags)
        for one, two in ((dense, hashed), (dense_b, hashed_b)):
            for k in one:
                if isinstance(one[k], np.ndarray):
                    assert np.array_equal(one[k], two[k], equal_nan=True), (cap, k)
    assert dense["kmer_breaks"][-1] > 4000


# ---- scaffold sets scored from their parts (bs_score_scaffolds) --------------------------------------

import scaffold_cases as SC  # noqa: E402
from test_scaffold_sets import SETS as SCAFFOLD_SETS  # noqa: E402


@pytest.mark.parametrize("kw", SCAFFOLD_SETS + [dict(seed=66, length=50000, read_len=150, coverage=30, n_base=16, n_scaffolds=60, overlap=30),
                                                dict(seed=67, length=50000, read_len=100, coverage=30, n_base=24, n_scaffolds=60, overlap=0, ragged=True)],
                         ids=lambda k: f"seed{k['seed']}")
@pytest.mark.parametrize("mode", ["scored_in_place", "weights", "global_rows", "small_hash", "ks_from_parts", "ks_from_parts_weights", "with_text", "junctions_probed"])
def test_scaffold_sets_vs_rescan_and_oracle(kw, mode, gpu_scorer, oracle, kmers, prob, monkeypatch):
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
    truth, reads, sset = SC.make_set(**kw)
    before = gpu_scorer.launch_count
    SC.check_scaffolds(gpu_scorer, oracle, kmers, prob, truth, reads, sset, flags=SC.mode_flags(mode))
    assert gpu_scorer.launch_count > before


@pytest.mark.parametrize("name,base,chains,reads,truth,kmer", SC.hand_sets(), ids=[h[0] for h in SC.hand_sets()])
@pytest.mark.parametrize("mode", ["scored_in_place", "weights", "global_rows", "small_hash", "ks_from_parts", "ks_from_parts_weights", "with_text", "junctions_probed"])
def test_scaffold_hand_built_sets(name, base, chains, reads, truth, kmer, mode, gpu_scorer, oracle, kmers, prob, monkeypatch):
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
    SC.check_scaffolds(gpu_scorer, oracle, kmers, prob, truth, reads, SC.hand_scaffold_set(base, chains), kmer=kmer, flags=SC.mode_flags(mode))


def test_scaffold_second_table_and_assembled_set(gpu_scorer, oracle, kmers, prob):
    truth, reads, sset = SC.make_set(65, length=20000, read_len=50, coverage=20, n_base=10, n_scaffolds=40, overlap=11)
    SC.check_scaffolds(gpu_scorer, oracle, kmers, prob, truth, reads, sset, second=tables.uniform(len(prob)))
    gpu_scorer.set_second_table(None)
    # the upstream sequence of calls: assemble_contigs -> calc_breakscore, here assemble_scaffolds -> score_scaffolds
    seg = synth.make_segment(77, length=20000, read_len=100, coverage=20, n_contigs=9, mut_frac=0.0)
    contigs = [seg.truth[s:s + len(c) + 60] for s, c in zip(seg.contig_truth_start, seg.contigs)]
    strings, aset = B.assemble_scaffolds(contigs, 13, 1234, n_shuffles=2000)
    got = SC.check_scaffolds(gpu_scorer, oracle, kmers, prob, seg.truth, seg.read_list, aset, oracle_sample=range(0, len(strings), max(1, len(strings) // 24)))
    assert got["sequence"] == strings and len(strings) >= 2


def test_cfg4_full_size_from_parts(gpu_scorer, kmers, prob):
    """BASELINE.json configs[3] at its size: the 10 000 scaffolds scored from their parts == the same 10 000 texts re-scanned
    (integer columns and KS-A bit for bit, fp64 sums to 1e-9), 64 of them against the CPU oracle with all reads, and 20
    repeats of the call bit-identical (shared-memory atomicMin rows, per-block scratch)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import loader as O
    seg = synth.make_scaffold_set(1400, n_scaffolds=10000)
    sset = B.ScaffoldSet(seg.base_contigs, seg.part_start, seg.part_base, np.zeros(len(seg.part_base), np.int32))
    gpu_scorer.set_table(kmers, prob)
    flags = B.WANT_KS | B.WANT_STARTPOS
    got = gpu_scorer.score_scaffolds(sset, seg.reads, seg.truth, flags=flags)
    ref = gpu_scorer.score(seg.contigs, seg.reads, seg.truth, flags=flags)
    assert got["sequence"] == seg.contigs
    for k in ("sequence_len", "kmer_breaks", "path_prob_dist_startpos", "ks_stat_prob_dist"):
        assert np.array_equal(got[k], ref[k], equal_nan=True), k
    for k in ("bp_score", "bp_score_norm_by_break_freqs", "bp_score_norm_by_len"):
        np.testing.assert_allclose(got[k], ref[k], rtol=1e-9, atol=0, err_msg=k)
    np.testing.assert_allclose(got["ks_stat_path_freq"], ref["ks_stat_path_freq"], rtol=1e-9, atol=1e-12, equal_nan=True)
    lens = np.array([len(c) for c in seg.contigs])
    chosen = sorted({int(np.argmax(lens)), int(np.argmin(lens))} | {int(i) for i in np.random.default_rng(45).permutation(len(lens))[:62]})
    reads = seg.read_list

    def check(idx):
        want = O.oracle_calc_breakscore([seg.contigs[i] for i in idx], reads, seg.truth, 8, kmers, prob, want_prob_dist=False)
        for k in ("sequence_len", "kmer_breaks", "path_prob_dist_startpos"):
            assert np.array_equal(got[k][idx], want[k]), k
        for k in ("bp_score", "bp_score_norm_by_break_freqs", "bp_score_norm_by_len"):
            np.testing.assert_allclose(got[k][idx], want[k], rtol=1e-9, atol=0, err_msg=k)
        for k in ("ks_stat_prob_dist", "ks_stat_path_freq"):
            np.testing.assert_allclose(got[k][idx], want[k], rtol=1e-9, atol=1e-12, equal_nan=True, err_msg=k)
        return len(idx)

    with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as ex:
        assert sum(ex.map(check, [chosen[i:i + 4] for i in range(0, len(chosen), 4)])) >= 62
    for it in range(20):
        again = gpu_scorer.score_scaffolds(sset, seg.reads, seg.truth, flags=flags)
        for k in got:
            if k != "sequence":
                assert np.array_equal(got[k], again[k], equal_nan=True), (it, k)


@pytest.mark.parametrize("mode", ["scored_in_place", "ks_from_parts"])
def test_scaffold_longer_than_65535_windows(mode, gpu_scorer, oracle, kmers, prob):
    from test_scaffold_sets import long_scaffold_case
    truth, reads, sset = long_scaffold_case()
    SC.check_scaffolds(gpu_scorer, oracle, kmers, prob, truth, reads, sset, flags=SC.mode_flags(mode) & ~B.WANT_POS & ~B.WANT_HIST)


@pytest.mark.parametrize("fallback", [False, True], ids=["from_parts", "texts_materialised"])
def test_scaffolds_device_resident(fallback, gpu_scorer, kmers, prob, monkeypatch):
    """bs_score_scaffolds with reads, truth and results on the device (what bench.py times), also through the fall-back for
    sets with too many (base contig, read) cells, against the host-buffer call of the same set"""
    import ctypes as C
    import torch
    if fallback:
        monkeypatch.setenv("BS_COMPOSE_MAX_CELLS", "1000")
    truth, reads, sset = SC.make_set(69, length=20000, read_len=100, coverage=20, n_base=10, n_scaffolds=50, overlap=12)
    gpu_scorer.set_table(kmers, prob)
    flags = B.WANT_KS | B.WANT_STARTPOS
    want = gpu_scorer.score_scaffolds(sset, reads, truth, flags=flags)
    n = len(sset)
    rd = torch.from_numpy(np.frombuffer(b"".join(reads), np.uint8).copy()).cuda()
    tr = torch.from_numpy(np.frombuffer(truth, np.uint8).copy()).cuda()
    i32 = torch.zeros(4, n, dtype=torch.int32, device="cuda")
    f64 = torch.zeros(5, n, dtype=torch.float64, device="cuda")
    r = B._Result()
    r.sequence_len, r.kmer_breaks, r.path_prob_dist_startpos, r.lev_dist_vs_true = [i32[i].data_ptr() for i in range(4)]
    (r.bp_score, r.bp_score_norm_by_break_freqs, r.bp_score_norm_by_len, r.ks_stat_prob_dist,
     r.ks_stat_path_freq) = [f64[i].data_ptr() for i in range(5)]
    st = sset.c_struct()
    for _ in range(2):
        gpu_scorer._check(gpu_scorer._lib.bs_score_scaffolds(gpu_scorer._ctx, C.byref(st), rd.data_ptr(), None, len(reads), 100, tr.data_ptr(),
                                                             len(truth), 8, flags | B.DEVICE_CHARS | B.DEVICE_RESULT, C.byref(r)))
    gpu_scorer.synchronize()
    torch.cuda.synchronize()
    for j, k in enumerate(("sequence_len", "kmer_breaks", "path_prob_dist_startpos")):
        assert np.array_equal(i32[j].cpu().numpy(), want[k]), k
    assert np.array_equal(f64[0].cpu().numpy(), want["bp_score"])
    assert np.array_equal(f64[3].cpu().numpy(), want["ks_stat_prob_dist"], equal_nan=True)
    assert np.array_equal(f64[4].cpu().numpy(), want["ks_stat_path_freq"], equal_nan=True)


def test_a_scaffolds_record_does_not_depend_on_the_rest_of_the_set(gpu_scorer, kmers, prob, monkeypatch):
    SC.check_set_independence(gpu_scorer, kmers, prob, monkeypatch, seed=72, length=50000, read_len=150, coverage=30, n_base=16,
                              n_scaffolds=80, overlap=20)
