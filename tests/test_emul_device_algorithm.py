"""The DEVICE ALGORITHM (the same kernel sources, compiled against tests/emul/cuda_emul.h) checked
against the oracle and the reference vectors on machines without a GPU.  This is not the product
path -- the GPU tests (test_gpu_parity.py) run the nvcc build through the same C-ABI."""
import numpy as np
import pytest

import parity_cases as P
from conftest import load_ref_vectors
from genomeassembler_dev_b200 import breakscore as B
from genomeassembler_dev_b200 import tables

CASES = [c for c in load_ref_vectors() if c["table"] in ("real", "rowid")]


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_reference_vectors(case, emul_scorer, kmers, table_set):
    P.check_reference_vector(emul_scorer, case, kmers, table_set)


@pytest.mark.parametrize("params", P.SMALL, ids=[f"L{p[1]}_r{p[2]}" for p in P.SMALL])
@pytest.mark.parametrize("mode", [0, B.PLACE_TILE, B.PLACE_SCAN], ids=["read_index", "tile_index", "scan"])
def test_small_segments_vs_oracle(params, mode, emul_scorer, oracle, kmers, prob):
    seg = P.make(*params)
    P.check_segment(emul_scorer, oracle, kmers, prob, seg, flags=P.FULL | mode)


@pytest.mark.parametrize("mode", [0, B.PLACE_TILE, B.PLACE_SCAN], ids=["read_index", "tile_index", "scan"])
@pytest.mark.parametrize("name,contigs,reads,truth,kmer", P.edge_inputs(), ids=[e[0] for e in P.edge_inputs()])
def test_edge_inputs(name, contigs, reads, truth, kmer, mode, emul_scorer, oracle, kmers, prob):
    from genomeassembler_dev_b200.synth import Segment
    seg = Segment(truth, None, contigs)
    P.check_segment(emul_scorer, oracle, kmers, prob, seg, kmer=kmer, reads=reads, flags=P.FULL | mode)


def test_random_pass_keeps_real_truth_table(emul_scorer, oracle, kmers, prob):
    """lib/DeNovoAssembler.R:326-333: uniform scoring table, real probabilities on the truth side"""
    seg = P.make(41, 2500, 50, 10, 5, 0)
    P.check_segment(emul_scorer, oracle, kmers, tables.uniform(len(prob)), seg, truth_prob=prob)


def test_set_table_drops_an_earlier_truth_table(emul_lib, oracle, kmers, prob):
    """A truth-side table is indexed by the rows of the table it was set against: bs_set_table on its own (direct
    C-ABI use, no bs_set_truth_table afterwards) must fall back to the new scoring table, whatever the row count."""
    import ctypes as C
    seg = P.make(43, 2000, 40, 8, 4, 0)
    sc = B.BreakageScorer(0, emul_lib)
    try:
        sc.set_table(kmers, tables.uniform(len(prob)), truth_prob=prob)  # truth-side table set against these rows
        perm = np.random.default_rng(5).permutation(len(kmers))           # same n, another key -> row mapping
        k2, p2 = [kmers[i] for i in perm], prob[perm]
        chars, off = B.flatten(k2)
        sc._check(sc._lib.bs_set_table(sc._ctx, C.c_void_p(chars.ctypes.data), C.c_void_p(off.ctypes.data),
                                       C.c_void_p(p2.ctypes.data), len(p2)))
        sc.n_table = len(p2)
        got = sc.score(seg.contigs, seg.reads, seg.truth, flags=B.DEFAULT_FLAGS)
        want = oracle.oracle_calc_breakscore(seg.contigs, seg.read_list, seg.truth, 8, k2, p2)
        for k in ("ks_stat_prob_dist", "ks_stat_path_freq"):
            np.testing.assert_allclose(got[k], want[k], rtol=1e-9, atol=1e-12, equal_nan=True, err_msg=k)
    finally:
        sc.close()


def test_batch_equals_single_calls(emul_scorer, kmers, prob):
    from genomeassembler_dev_b200 import synth
    emul_scorer.set_table(kmers, prob)
    b = synth.make_batch(3, seed=50, length=2000, read_len=40, coverage=8, contigs_lo=2, contigs_hi=5)
    res = emul_scorer.score_batch(b.read_chars, None, b.read_len, b.contig_chars, b.contig_off, b.truth_chars,
                                  b.truth_off, b.seg_read_start, b.seg_contig_start, flags=B.DEFAULT_FLAGS | B.WANT_HIST)
    for s in range(b.n_segments):
        seg = b.segment(s)
        one = emul_scorer.score(seg.contigs, seg.reads, seg.truth, flags=B.DEFAULT_FLAGS | B.WANT_HIST)
        c0, c1 = int(b.seg_contig_start[s]), int(b.seg_contig_start[s + 1])
        for k in ("sequence_len", "kmer_breaks", "path_prob_dist_startpos", "bp_score", "bp_score_norm_by_break_freqs",
                  "ks_stat_prob_dist", "ks_stat_path_freq", "hist"):
            assert np.array_equal(res[k][c0:c1], one[k], equal_nan=True), k


def test_argument_errors(emul_scorer, kmers, prob):
    with pytest.raises(B.BreakscoreError):
        emul_scorer.set_table(["ACGX"], [0.5])
    with pytest.raises(B.BreakscoreError):
        emul_scorer.set_table(["ACGTACGTA"], [0.5])
    emul_scorer.set_table(kmers, prob)
    with pytest.raises(B.BreakscoreError):
        emul_scorer.score([b"ACGT"], [b"AC"], b"ACGT", kmer=0)
    with pytest.raises(B.BreakscoreError):
        emul_scorer.score_batch(np.zeros(4, np.uint8), None, 2, np.zeros(4, np.uint8), [0, 4, 2], np.zeros(4, np.uint8),
                                [0, 4], [0, 2], [0, 2])


def test_chunked_pipeline_equals_one_chunk(emul_lib, emul_scorer, kmers, prob, monkeypatch):
    """many small pipeline chunks (two workspaces reused) == one chunk, bit for bit"""
    from genomeassembler_dev_b200 import synth
    b = synth.make_batch(7, seed=90, length=1500, read_len=30, coverage=6, contigs_lo=1, contigs_hi=4)
    args = (b.read_chars, None, b.read_len, b.contig_chars, b.contig_off, b.truth_chars, b.truth_off,
            b.seg_read_start, b.seg_contig_start)
    flags = B.DEFAULT_FLAGS | B.WANT_HIST | B.WANT_POS
    emul_scorer.set_table(kmers, prob)
    one = emul_scorer.score_batch(*args, flags=flags)
    monkeypatch.setenv("BS_CHUNK_KB", "12")
    with B.BreakageScorer(0, emul_lib) as sc:
        sc.set_table(kmers, prob)
        many = sc.score_batch(*args, flags=flags)
    for k in one:
        assert np.array_equal(one[k], many[k], equal_nan=True), k


def test_poll_callback_interrupts_between_chunks(emul_lib, kmers, prob, monkeypatch):
    """bs_ctx_set_poll: called between pipeline chunks on the calling thread; a true return ends the call with
    BS_ERR_INTERRUPTED, the context stays usable and the next full call is bit-identical to an undisturbed one."""
    from genomeassembler_dev_b200 import synth
    b = synth.make_batch(7, seed=91, length=1500, read_len=30, coverage=6, contigs_lo=1, contigs_hi=4)
    args = (b.read_chars, None, b.read_len, b.contig_chars, b.contig_off, b.truth_chars, b.truth_off,
            b.seg_read_start, b.seg_contig_start)
    monkeypatch.setenv("BS_CHUNK_KB", "12")
    with B.BreakageScorer(0, emul_lib) as sc:
        sc.set_table(kmers, prob)
        want = sc.score_batch(*args)
        calls = []
        sc.set_poll(lambda: calls.append(1) or False)  # never interrupts: same results, polled once per later chunk
        got = sc.score_batch(*args)
        assert len(calls) >= 2
        for k in want:
            assert np.array_equal(want[k], got[k], equal_nan=True), k
        calls.clear()
        sc.set_poll(lambda: calls.append(1) or len(calls) >= 2)  # interrupt at the second poll
        with pytest.raises(B.BreakscoreError) as ei:
            sc.score_batch(*args)
        assert ei.value.code == B.ERR_INTERRUPTED and len(calls) == 2
        sc.set_poll(None)
        again = sc.score_batch(*args)
        for k in want:
            assert np.array_equal(want[k], again[k], equal_nan=True), k


def test_one_segment_over_several_contexts(emul_lib, emul_scorer, kmers, prob):
    """bs_score_multi: the contigs of one segment dealt out over three contexts (one host thread each, reads replicated)
    == one call on one context, byte for byte, every output including the variable-length ones."""
    seg = P.make(*P.SMALL[1], mut=0.3)
    contigs = list(seg.contigs) + [seg.contigs[0][:5], b"", seg.contigs[-1]]  # short, empty and duplicate contigs too
    flags = B.DEFAULT_FLAGS | B.WANT_HIST | B.WANT_POS | B.WANT_LEV | B.WANT_SECOND_TABLE
    emul_scorer.set_table(kmers, prob)
    emul_scorer.set_second_table(tables.uniform(len(prob)))
    want = emul_scorer.score(contigs, seg.reads, seg.truth, flags=flags)
    others = [B.BreakageScorer(0, emul_lib) for _ in range(2)]
    try:
        for o in others:
            o.set_table(kmers, prob)
            o.set_second_table(tables.uniform(len(prob)))
        got = emul_scorer.score(contigs, seg.reads, seg.truth, flags=flags, group=others)
        more = emul_scorer.score(contigs[:2], seg.reads, seg.truth, flags=flags, group=others)  # fewer contigs than contexts
        two = emul_scorer.score(contigs[:2], seg.reads, seg.truth, flags=flags)
        with pytest.raises(B.BreakscoreError, match="same as context"):  # one context twice would be driven by two threads
            emul_scorer.score(contigs, seg.reads, seg.truth, flags=flags, group=[others[0], others[0]])
        others[1].set_table(kmers[:16], np.full(16, 1 / 16))  # a context with another table is refused
        with pytest.raises(B.BreakscoreError):
            emul_scorer.score(contigs, seg.reads, seg.truth, flags=flags, group=others)
    finally:
        for o in others:
            o.close()
    for a, b in ((want, got), (two, more)):
        for k in a:
            if k in ("path_prob_dist", "path_prob_dist2"):
                assert all(np.array_equal(x, y) for x, y in zip(a[k], b[k])), k
            elif k != "sequence":
                assert np.array_equal(a[k], b[k], equal_nan=True), k


# ---- infix edit distance (lev_dist_vs_true) ---------------------------------------------------

@pytest.mark.parametrize("params", P.SMALL[:4], ids=[f"L{p[1]}_r{p[2]}" for p in P.SMALL[:4]])
def test_edit_distance_small_segments(params, emul_scorer, oracle, kmers, prob):
    seg = P.make(*params, mut=0.6)
    P.check_segment(emul_scorer, oracle, kmers, prob, seg, flags=P.FULL | B.WANT_LEV)


@pytest.mark.parametrize("name,contigs,reads,truth,kmer", P.edge_inputs(), ids=[e[0] for e in P.edge_inputs()])
def test_edit_distance_edge_inputs(name, contigs, reads, truth, kmer, emul_scorer, oracle, kmers, prob):
    from genomeassembler_dev_b200.synth import Segment
    P.check_segment(emul_scorer, oracle, kmers, prob, Segment(truth, None, contigs), kmer=kmer, reads=reads,
                    flags=B.DEFAULT_FLAGS | B.WANT_LEV)


def test_edit_distance_indels_and_long_contigs(emul_scorer, oracle, kmers, prob):
    """contigs longer than one warp of pattern blocks (2048 bases), with substitutions, insertions,
    deletions and unrelated sequence"""
    from genomeassembler_dev_b200 import synth
    rng = np.random.default_rng(9)
    truth = synth.codes_to_ascii(synth.random_truth_codes(rng, 5200)).tobytes()
    a = bytearray(truth[100:2700]); a[1300] = ord("A") if a[1300] != ord("A") else ord("C")   # 2600 bases, 1 substitution
    b = truth[300:1500] + truth[1510:2900]                                                    # 10-base deletion
    c = truth[2000:3000] + b"ACGTTGCA" + truth[3000:4700]                                     # 8-base insertion, 2708 bases
    d = synth.codes_to_ascii(synth.random_truth_codes(rng, 300)).tobytes()                    # unrelated
    e = truth[4000:5200] + b"GGGGGGGGGG"                                                      # overhang at the end of the truth
    seg = synth.Segment(truth, None, [bytes(a), b, c, d, e, truth[50:2500]])
    reads = [truth[i:i + 40] for i in range(0, 5000, 97)]
    got, want = P.check_segment(emul_scorer, oracle, kmers, prob, seg, reads=reads, flags=B.DEFAULT_FLAGS | B.WANT_LEV)
    assert got["lev_dist_vs_true"][0] == 1 and got["lev_dist_vs_true"][5] == 0


def check_second_table(scorer, kmers, prob, seg):
    """both table passes of the R driver (real, then uniform with the real table on the truth side) in ONE
    call == two calls, bit for bit"""
    uni = tables.uniform(len(prob))
    flags = B.DEFAULT_FLAGS
    scorer.set_table(kmers, prob)
    scorer.set_second_table(uni)
    both = scorer.score(seg.contigs, seg.read_list, seg.truth, flags=flags | B.WANT_SECOND_TABLE)
    first = scorer.score(seg.contigs, seg.read_list, seg.truth, flags=flags)
    scorer.set_table(kmers, uni, truth_prob=prob)
    second = scorer.score(seg.contigs, seg.read_list, seg.truth, flags=flags)
    for k in ("bp_score", "bp_score_norm_by_break_freqs", "bp_score_norm_by_len", "ks_stat_prob_dist", "ks_stat_path_freq",
              "kmer_breaks", "path_prob_dist_startpos"):
        assert np.array_equal(both[k], first[k], equal_nan=True), k
    for k in ("bp_score", "bp_score_norm_by_break_freqs", "bp_score_norm_by_len", "ks_stat_prob_dist", "ks_stat_path_freq"):
        assert np.array_equal(both[k + "2"], second[k], equal_nan=True), k + "2"
    for a, b in zip(both["path_prob_dist2"], second["path_prob_dist"]):
        assert np.array_equal(a, b)
    scorer.set_table(kmers, prob)
    with pytest.raises(B.BreakscoreError):  # bs_set_table removes the second table
        scorer.score(seg.contigs, seg.read_list, seg.truth, flags=flags | B.WANT_SECOND_TABLE)


def test_second_table_in_one_call(emul_scorer, kmers, prob):
    check_second_table(emul_scorer, kmers, prob, P.make(43, 2500, 50, 10, 5, 1))


def check_irregular_batch(scorer, oracle, kmers, prob, chunk_env=None):
    """segments without contigs, without reads, with ragged reads and with very long reads, in one batch"""
    from genomeassembler_dev_b200 import synth
    rng = np.random.default_rng(77)
    segs = []
    for i, (L, r, nc) in enumerate([(1500, 30, 3), (1200, 25, 0), (1800, 40, 4), (1000, 20, 2), (3000, 1000, 2)]):
        segs.append(synth.make_segment(500 + i, length=L, read_len=r, coverage=6 if r < 100 else 3, n_contigs=max(nc, 1), mut_frac=0.3))
    reads, contigs, truths, srs, scs = [], [], [], [0], [0]
    for i, sg in enumerate(segs):
        rl = sg.read_list
        if i == 0:
            rl = [x[: int(rng.integers(5, 31))] for x in rl]   # ragged
        if i == 3:
            rl = []                                            # no reads
        cs = [] if i == 1 else sg.contigs                      # no contigs
        reads += rl; contigs += cs; truths.append(sg.truth)
        srs.append(len(reads)); scs.append(len(contigs))
    rd, rd_off = B.flatten(reads)
    ct, ct_off = B.flatten(contigs)
    tr, tr_off = B.flatten(truths)
    scorer.set_table(kmers, prob)
    res = scorer.score_batch(rd, rd_off, 0, ct, ct_off, tr, tr_off, srs, scs, flags=B.DEFAULT_FLAGS | B.WANT_HIST | B.WANT_LEV)
    for s in range(len(segs)):
        c0, c1 = scs[s], scs[s + 1]
        if c1 == c0:
            continue
        want = oracle.oracle_calc_breakscore(contigs[c0:c1], reads[srs[s]:srs[s + 1]], truths[s], 8, kmers, prob,
                                             want_hist=True, want_lev=True)
        for k in ("sequence_len", "kmer_breaks", "path_prob_dist_startpos", "hist", "lev_dist_vs_true"):
            assert np.array_equal(res[k][c0:c1], want[k]), (s, k)
        for k in ("bp_score", "bp_score_norm_by_break_freqs", "ks_stat_prob_dist", "ks_stat_path_freq"):
            np.testing.assert_allclose(res[k][c0:c1], want[k], rtol=1e-9, atol=1e-12, equal_nan=True, err_msg=f"{s} {k}")
        off = res["path_prob_dist_off"]
        for c in range(c0, c1):
            assert np.array_equal(res["path_prob_dist_flat"][off[c]:off[c + 1]], want["path_prob_dist"][c - c0])
    return res


def test_irregular_batch(emul_scorer, oracle, kmers, prob):
    check_irregular_batch(emul_scorer, oracle, kmers, prob)


def test_irregular_batch_in_small_chunks(emul_lib, oracle, kmers, prob, monkeypatch):
    monkeypatch.setenv("BS_CHUNK_KB", "3")
    with B.BreakageScorer(0, emul_lib) as sc:
        check_irregular_batch(sc, oracle, kmers, prob)


def check_uniform_read_lengths(scorer, oracle, kmers, prob, lengths):
    """dense reads of one length (the fast packing kernel): word-boundary lengths, tiles of packed words that
    span several small segments, reads with bytes outside ACGT"""
    rng = np.random.default_rng(99)
    scorer.set_table(kmers, prob)
    for L in lengths:
        truths, contigs, reads, srs, scs = [], [], [], [0], [0]
        for s, n_reads in enumerate([3, 1, 0, 9, 2, 14]):
            t = bytes(rng.choice(list(b"ACGT"), size=L + 60 + 17 * s).astype(np.uint8))
            cs = [t[5:], t[: L + 20], t[3:L + 3] + b"N" + t[L + 4:L + 30]]
            rl = []
            for i in range(n_reads):
                a = int(rng.integers(0, len(t) - L + 1))
                x = bytearray(t[a:a + L])
                if i % 5 == 4:
                    x[int(rng.integers(0, L))] = ord("N")      # flagged: placed by byte comparison
                if i % 7 == 6:
                    x[L - 1] = ord("acgt"[i % 4])               # lowercase never matches
                rl.append(bytes(x))
            truths.append(t); contigs += cs; reads += rl
            srs.append(len(reads)); scs.append(len(contigs))
        rd, _ = B.flatten(reads)
        ct, ct_off = B.flatten(contigs)
        tr, tr_off = B.flatten(truths)
        res = scorer.score_batch(rd, None, L, ct, ct_off, tr, tr_off, srs, scs, flags=B.DEFAULT_FLAGS | B.WANT_HIST | B.WANT_POS)
        for s in range(len(truths)):
            c0, c1 = scs[s], scs[s + 1]
            want = oracle.oracle_calc_breakscore(contigs[c0:c1], reads[srs[s]:srs[s + 1]], truths[s], 8, kmers, prob,
                                                 want_hist=True, want_pos=True)
            for k in ("kmer_breaks", "path_prob_dist_startpos", "hist"):
                assert np.array_equal(res[k][c0:c1], want[k]), (L, s, k)
            np.testing.assert_allclose(res["bp_score"][c0:c1], want["bp_score"], rtol=1e-9, err_msg=f"L={L} seg={s}")


UNIFORM_LENGTHS = [1, 5, 16, 31, 32, 33, 64, 65, 96, 127, 128, 129, 150, 160, 192, 321]


def test_uniform_read_lengths(emul_scorer, oracle, kmers, prob):
    check_uniform_read_lengths(emul_scorer, oracle, kmers, prob, UNIFORM_LENGTHS)


def check_two_phase(scorer, oracle, kmers, prob, seg, n_shards=3):
    """BS_WEIGHTS_OUT / BS_WEIGHTS_IN: the reads of one segment in n disjoint shards placed one shard at a time, the
    position weights summed (what the all-reduce of sharding.score_reads_sharded does), the contigs scored from the sum
    in two separate ranges == one call with all the reads, every output bit for bit; and == the oracle."""
    scorer.set_table(kmers, prob)
    flags = B.DEFAULT_FLAGS | B.WANT_HIST
    want = scorer.score(seg.contigs, seg.reads, seg.truth, flags=flags)
    lens = np.array([len(c) for c in seg.contigs])
    off = np.concatenate([[0], np.cumsum(lens)])
    C = len(seg.contigs)
    w_sum, t_sum = np.zeros(int(off[-1]) + C + 1, np.int32), np.zeros(C + 1, np.int32)
    for k in range(n_shards):
        w, t = np.full_like(w_sum, -7), np.full_like(t_sum, -7)  # (the library must overwrite, not accumulate)
        scorer.place_weights(seg.contigs, seg.reads[k::n_shards], w.ctypes.data, t.ctypes.data)
        scorer.synchronize()
        w_sum[:-1] += w[:-1]
        t_sum[:-1] += t[:-1]
    assert np.array_equal(t_sum[:C], want["kmer_breaks"])
    cut = C // 2
    parts = []
    for c0, c1 in ((0, cut), (cut, C)):
        parts.append(scorer.score_from_weights(seg.contigs[c0:c1], seg.truth, w_sum.ctypes.data + 4 * (int(off[c0]) + c0),
                                               t_sum.ctypes.data + 4 * c0, flags=flags))
    for k in want:
        if k == "sequence":
            continue
        if k == "path_prob_dist":
            got = parts[0][k] + parts[1][k]
            assert all(np.array_equal(a, b) for a, b in zip(got, want[k])), k
        else:
            assert np.array_equal(np.concatenate([parts[0][k], parts[1][k]]), want[k], equal_nan=True), k
    ref = oracle.oracle_calc_breakscore(seg.contigs, seg.read_list, seg.truth, 8, kmers, prob, want_hist=True)
    assert np.array_equal(want["hist"], ref["hist"])
    with pytest.raises(B.BreakscoreError):  # the two flags exclude each other; weights pointers are required
        scorer.score_batch(np.zeros(1, np.uint8), None, 1, *B.flatten(seg.contigs), *B.flatten([seg.truth]), [0, 0], [0, C], flags=B.WEIGHTS_IN)


def test_two_phase_scoring(emul_scorer, oracle, kmers, prob):
    check_two_phase(emul_scorer, oracle, kmers, prob, P.make(*P.SMALL[0], mut=0.3))
    check_two_phase(emul_scorer, oracle, kmers, prob, P.make(*P.SMALL[2], mut=0.3), n_shards=2)


def check_fused_scoring(scorer, oracle, kmers, prob, monkeypatch, long_len):
    """kmer == 8 without the dense histogram: contigs of at least FUSE_MIN_LEN bases (16384; 2560 under the emulation) are
    scored by the long-contig KS-A kernel on its way over the windows (position p <-> window p - 4; positions 0..3 and
    L-3..L-1 by the generic rule), the others by k_break_score.  Against the oracle, and against k_break_score for every
    contig (BS_FUSE_SCORE=0): integer outputs and KS bit-exact, the sums to 1e-12 (another summation order).  A contig's
    results do not depend on what else is in the call."""
    from genomeassembler_dev_b200 import synth
    rng = np.random.default_rng(808)
    Lt = 3 * long_len
    truth = synth.codes_to_ascii(synth.random_truth_codes(rng, Lt))
    starts = rng.integers(0, Lt - 60, size=Lt // 4)
    reads = [truth[a:a + 60].tobytes() for a in starts] + [truth[:3].tobytes(), truth[long_len - 2:long_len].tobytes(), truth[1:3].tobytes()]
    long_a, long_b = truth[:long_len].tobytes(), truth[long_len // 2:long_len // 2 + long_len + 77].tobytes()
    mutated = bytearray(truth[Lt - long_len - 5:Lt]); mutated[long_len // 3] ^= 6
    with_n = bytearray(truth[100:100 + long_len]); with_n[4000 % long_len] = ord("N"); with_n[5] = ord("N")
    shorts = [truth[50:50 + long_len - 1].tobytes(), truth[7:900].tobytes(), truth[300:309].tobytes(), b"ACGT"]
    scorer.set_table(kmers, prob)
    scorer.set_second_table(tables.uniform(len(prob)))
    flags = B.DEFAULT_FLAGS | B.WANT_SECOND_TABLE
    tr = truth.tobytes()
    for contigs, all_long in (([long_a, long_b, bytes(mutated), bytes(with_n)], True), ([long_a] + shorts + [long_b, bytes(with_n)], False)):
        monkeypatch.delenv("BS_FUSE_SCORE", raising=False)
        n0 = scorer.launch_count
        got = scorer.score(contigs, reads, tr, flags=flags)
        n_fused = scorer.launch_count - n0
        monkeypatch.setenv("BS_FUSE_SCORE", "0")
        n0 = scorer.launch_count
        ref = scorer.score(contigs, reads, tr, flags=flags)
        assert scorer.launch_count - n0 == n_fused + (2 if all_long else 0)  # no k_break_score launch when every contig is long
        monkeypatch.delenv("BS_FUSE_SCORE")
        want = oracle.oracle_calc_breakscore(contigs, reads, tr, 8, kmers, prob)
        assert want["kmer_breaks"].sum() > Lt // 8
        for k in ("sequence_len", "kmer_breaks", "path_prob_dist_startpos"):
            assert np.array_equal(got[k], ref[k]) and np.array_equal(got[k], want[k]), k
        for k in ("ks_stat_path_freq", "ks_stat_path_freq2", "ks_stat_prob_dist", "ks_stat_prob_dist2"):
            assert np.array_equal(got[k], ref[k], equal_nan=True), k
        for k in ("bp_score", "bp_score_norm_by_break_freqs", "bp_score_norm_by_len", "bp_score2", "bp_score_norm_by_break_freqs2"):
            np.testing.assert_allclose(got[k], ref[k], rtol=1e-12, atol=0, err_msg=k)
        for k in ("bp_score", "bp_score_norm_by_break_freqs", "bp_score_norm_by_len"):
            np.testing.assert_allclose(got[k], want[k], rtol=1e-9, atol=0, err_msg=k)
        for k in ("ks_stat_prob_dist", "ks_stat_path_freq"):
            np.testing.assert_allclose(got[k], want[k], rtol=1e-9, atol=1e-12, equal_nan=True, err_msg=k)
        for a, b in zip(got["path_prob_dist"], want["path_prob_dist"]):
            assert np.array_equal(a, b)
        if all_long:
            first = got
        else:  # the same contig in another call: the same bits
            for k in ("bp_score", "bp_score_norm_by_break_freqs", "ks_stat_path_freq", "ks_stat_prob_dist", "bp_score2"):
                assert got[k][0] == first[k][0] and got[k][-2] == first[k][1] and (got[k][-1] == first[k][3] or np.isnan(got[k][-1])), k
    scorer.set_second_table(None)


def test_fused_scoring(emul_scorer, oracle, kmers, prob, monkeypatch):
    check_fused_scoring(emul_scorer, oracle, kmers, prob, monkeypatch, long_len=2600)


def check_pack_variants(scorer, kmers, prob, monkeypatch, lengths, to_dev=None, n_reads=(1, 3, 257, 1111)):
    """Reads of one length through both packing kernels (bulk-copy staged: default; register staged: BS_PACK_BULK=0)
    from read buffers that start at every 16-byte phase (BS_DEVICE_CHARS: the library packs straight out of the
    caller's buffer) and end anywhere, with bytes outside ACGT inside, at the ends and next to tile boundaries:
    positions, flags-dependent placements and break counts identical to the aligned host-buffer call."""
    import ctypes as C
    from genomeassembler_dev_b200 import synth
    rng = np.random.default_rng(77)
    scorer.set_table(kmers, prob)
    for L in lengths:
        for N in n_reads:
            Lt = max(4 * L, 600)
            truth = synth.codes_to_ascii(synth.random_truth_codes(rng, Lt))
            starts = rng.integers(0, Lt - L + 1, size=N)
            reads = truth[starts[:, None] + np.arange(L)[None, :]].copy()
            for i in rng.integers(0, N, size=max(1, N // 40)):  # a few reads with a byte outside ACGT (first, last, middle)
                reads[i, int(rng.choice([0, L - 1, L // 2]))] = ord("N")
            tr = truth.copy()
            tr[rng.integers(0, Lt, size=3)] = ord("N")
            contigs = [tr[:Lt // 2].tobytes(), tr[Lt // 3:].tobytes(), tr.tobytes()]
            ct, ct_off = B.flatten(contigs)
            trc, tr_off = B.flatten([tr.tobytes()])
            flags = B.WANT_POS | B.WANT_STARTPOS
            want = scorer.score_batch(reads.reshape(-1), None, L, ct, ct_off, trc, tr_off, [0, N], [0, 3], flags=flags)
            for phase in (0, 1, 7, 15):
                for bulk in ("1", "0"):
                    monkeypatch.setenv("BS_PACK_BULK", bulk)
                    buf = np.full(phase + N * L, ord("T"), np.uint8)  # exact size: nothing behind the last read
                    buf[phase:] = reads.reshape(-1)
                    keep = [buf, ct, trc]
                    if to_dev is not None:
                        dbuf, dct, dtr = to_dev(buf), to_dev(ct), to_dev(trc)
                        keep += [dbuf, dct, dtr]
                        rp, cp, tp = dbuf.data_ptr() + phase, dct.data_ptr(), dtr.data_ptr()
                    else:
                        rp, cp, tp = buf.ctypes.data + phase, ct.ctypes.data, trc.ctypes.data
                    srs, scs = np.array([0, N], np.int64), np.array([0, 3], np.int64)
                    bt = B._Batch(1, N, 3, rp, None, L, cp, ct_off.ctypes.data, tp, tr_off.ctypes.data, srs.ctypes.data, scs.ctypes.data)
                    out_i = np.zeros((3, 3), np.int32)
                    pos = np.zeros(3 * N, np.int32)
                    pos_off = np.arange(4, dtype=np.int64) * N
                    r = B._Result()
                    r.sequence_len, r.kmer_breaks, r.path_prob_dist_startpos = [out_i[i].ctypes.data for i in range(3)]
                    r.pos, r.pos_off = pos.ctypes.data, pos_off.ctypes.data
                    scorer.score_batch_raw(bt, r, 8, flags | B.DEVICE_CHARS)
                    assert np.array_equal(out_i[1], want["kmer_breaks"]), (L, N, phase, bulk)
                    assert np.array_equal(out_i[2], want["path_prob_dist_startpos"]), (L, N, phase, bulk)
                    assert np.array_equal(pos, want["pos_flat"][:3 * N]), (L, N, phase, bulk)
                    del keep


def test_pack_variants(emul_scorer, kmers, prob, monkeypatch):
    check_pack_variants(emul_scorer, kmers, prob, monkeypatch, [1, 12, 150, 151, 1000], n_reads=(1, 3, 300))


def check_spectrum_variants(scorer, lib, kmers, prob):
    """truth spectrum with the rank table in shared memory (default) == with the table in global memory"""
    import os
    from genomeassembler_dev_b200 import synth
    b = synth.make_batch(9, seed=91, length=1800, read_len=30, coverage=5, contigs_lo=1, contigs_hi=4)
    args = (b.read_chars, None, b.read_len, b.contig_chars, b.contig_off, b.truth_chars, b.truth_off,
            b.seg_read_start, b.seg_contig_start)
    scorer.set_table(kmers, prob)
    one = scorer.score_batch(*args, flags=B.DEFAULT_FLAGS)
    os.environ["BS_SPECTRUM_TABLE"] = "0"
    try:
        with B.BreakageScorer(0, lib) as sc:
            sc.set_table(kmers, prob)
            two = sc.score_batch(*args, flags=B.DEFAULT_FLAGS)
    finally:
        del os.environ["BS_SPECTRUM_TABLE"]
    assert np.isfinite(one["ks_stat_prob_dist"]).any()
    for k in one:
        assert np.array_equal(one[k], two[k], equal_nan=True), k


def test_spectrum_variants(emul_scorer, emul_lib, kmers, prob):
    check_spectrum_variants(emul_scorer, emul_lib, kmers, prob)


# ---- hashed placement scratch (one huge segment, cfg-5) ---------------------------------------

def check_hashed_scratch(scorer, oracle, kmers, prob, seg, monkeypatch, block_threads):
    """the open-addressed (read -> leftmost position) table that replaces the dense per-block row when a
    segment has too many reads for one row per resident block: same results as the oracle and as the
    dense row, also when a contig overflows the first table and the launch is repeated with a larger one"""
    dense, _ = P.check_segment(scorer, oracle, kmers, prob, seg)
    n0 = scorer.launch_count
    scorer.score(seg.contigs, seg.read_list, seg.truth, flags=P.FULL)
    dense_launches = scorer.launch_count - n0
    monkeypatch.setenv("BS_PLACE_SCRATCH_MB", "0")
    hashed, _ = P.check_segment(scorer, oracle, kmers, prob, seg)
    monkeypatch.setenv("BS_PLACE_HASH_CAP", "1")   # smallest table: long contigs overflow it
    n0 = scorer.launch_count
    small, _ = P.check_segment(scorer, oracle, kmers, prob, seg)
    small_launches = scorer.launch_count - n0
    for k in dense:
        if isinstance(dense[k], np.ndarray):
            assert np.array_equal(dense[k], hashed[k], equal_nan=True), k
            assert np.array_equal(dense[k], small[k], equal_nan=True), k
    # smallest table: 4 slots per thread of a block, given up when half full (a few more reads can slip in)
    most = int(np.max(dense["kmer_breaks"]))
    if most >= 3 * block_threads:
        assert small_launches > dense_launches   # the placement kernel ran more than once
    elif most < 2 * block_threads:
        assert small_launches == dense_launches


@pytest.mark.parametrize("params", P.SMALL, ids=[f"L{p[1]}_r{p[2]}" for p in P.SMALL])
def test_hashed_scratch_small_segments(params, emul_scorer, oracle, kmers, prob, monkeypatch):
    seg = P.make(*params)
    check_hashed_scratch(emul_scorer, oracle, kmers, prob, seg, monkeypatch, block_threads=64)


@pytest.mark.parametrize("name,contigs,reads,truth,kmer", P.edge_inputs(), ids=[e[0] for e in P.edge_inputs()])
def test_hashed_scratch_edge_inputs(name, contigs, reads, truth, kmer, emul_scorer, oracle, kmers, prob, monkeypatch):
    from genomeassembler_dev_b200.synth import Segment
    monkeypatch.setenv("BS_PLACE_SCRATCH_MB", "0")
    monkeypatch.setenv("BS_PLACE_HASH_CAP", "1")
    P.check_segment(emul_scorer, oracle, kmers, prob, Segment(truth, None, contigs), kmer=kmer, reads=reads)


def test_hashed_scratch_duplicate_reads(emul_scorer, oracle, kmers, prob, monkeypatch):
    """thousands of copies of a few reads: every copy is its own table entry (reads are not deduplicated), so the
    table grows until it holds them all"""
    from genomeassembler_dev_b200 import synth
    rng = np.random.default_rng(5)
    truth = synth.codes_to_ascii(synth.random_truth_codes(rng, 900)).tobytes()
    reads = [truth[100:130]] * 700 + [truth[400:430]] * 500 + [truth[10:40]] * 3 + [b"ACGTNACGTAACCGGTTACGTACGTAAGGT"] * 4
    seg = synth.Segment(truth, None, [truth[50:600], truth[380:880], truth[100:130] + truth[100:130], truth[5:60]])
    monkeypatch.setenv("BS_PLACE_SCRATCH_MB", "0")
    monkeypatch.setenv("BS_PLACE_HASH_CAP", "1")
    n0 = emul_scorer.launch_count
    got, want = P.check_segment(emul_scorer, oracle, kmers, prob, seg, reads=reads)
    assert list(got["kmer_breaks"]) == [1200, 500, 700, 3]
    monkeypatch.delenv("BS_PLACE_SCRATCH_MB")
    n1 = emul_scorer.launch_count
    P.check_segment(emul_scorer, oracle, kmers, prob, seg, reads=reads)
    assert n1 - n0 >= (emul_scorer.launch_count - n1) + 3   # 256 -> 512 -> 1024 -> 2048 -> 4096 slots


# ---- contig-in-truth offsets with several seed-table groups per segment -------------------------

def check_startpos_many_contigs(scorer, oracle, kmers, prob, n_contigs=2500, L=30000):
    """more contigs in a segment than one seed table holds (groups of 1024), duplicates, contigs sharing their
    first 32 bases, mutated copies, contigs shorter than a seed; a second small segment in the same batch"""
    from genomeassembler_dev_b200 import synth
    rng = np.random.default_rng(2500)
    truths, contigs, reads, srs, scs = [], [], [], [0], [0]
    for s, (Ls, nc) in enumerate([(L, n_contigs), (3000, 7)]):
        t = synth.codes_to_ascii(synth.random_truth_codes(rng, Ls)).tobytes()
        cs = []
        for i in range(nc):
            a = int(rng.integers(0, Ls - 200))
            ln = int(rng.integers(20 if i % 50 == 0 else 33, 160))
            c = bytearray(t[a:a + ln])
            if i % 9 == 0 and ln > 40:
                c[int(rng.integers(32, ln))] ^= 6          # substitution after the seed: same seed, not a substring
            cs.append(bytes(c))
        cs += [cs[0], cs[1][:40], t[100:164], t[100:150]]   # duplicates and shared seeds
        rl = [t[a:a + 30] for a in range(0, Ls - 30, 7)]
        truths.append(t); contigs += cs; reads += rl
        srs.append(len(reads)); scs.append(len(contigs))
    rd, _ = B.flatten(reads)
    ct, ct_off = B.flatten(contigs)
    tr, tr_off = B.flatten(truths)
    scorer.set_table(kmers, prob)
    res = scorer.score_batch(rd, None, 30, ct, ct_off, tr, tr_off, srs, scs, flags=B.WANT_STARTPOS)
    for s in range(len(truths)):
        c0, c1 = scs[s], scs[s + 1]
        want = np.array([truths[s].find(c) for c in contigs[c0:c1]], dtype=np.int32) * (res["kmer_breaks"][c0:c1] > 0)
        assert np.array_equal(res["path_prob_dist_startpos"][c0:c1], want), s
    assert np.count_nonzero(res["path_prob_dist_startpos"] == -1) > 100
    return res


def test_startpos_many_contigs(emul_scorer, oracle, kmers, prob):
    check_startpos_many_contigs(emul_scorer, oracle, kmers, prob)


def check_startpos_geometries(scorer, oracle, kmers, prob, monkeypatch, **kw):
    """Both seed-table geometries of the contig-in-truth search on the same inputs (BS_STARTPOS_BIG forces one:
    groups of 1024 contigs with the table in shared memory / one table per segment in global memory behind a 10-base
    prefix bitmap), and a candidate queue of 5 entries (every further candidate is verified inside the scan)."""
    for env in ({"BS_STARTPOS_BIG": "0"}, {"BS_STARTPOS_BIG": "1"}, {"BS_STARTPOS_QCAP": "5"}, {"BS_STARTPOS_BIG": "0", "BS_STARTPOS_QCAP": "5"}):
        for k in ("BS_STARTPOS_BIG", "BS_STARTPOS_QCAP"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        check_startpos_many_contigs(scorer, oracle, kmers, prob, **kw)
        scorer.set_table(kmers, prob)
        for name, contigs, reads, truth, kmer in (P.edge_inputs() if "BS_STARTPOS_BIG" in env else []):
            got = scorer.score(contigs, reads, truth, kmer=kmer, flags=B.WANT_STARTPOS | B.WANT_LEV)
            want = oracle.oracle_calc_breakscore(contigs, reads, truth, kmer, kmers, prob, want_ks=False, want_lev=True, want_prob_dist=False)
            assert np.array_equal(got["path_prob_dist_startpos"], want["path_prob_dist_startpos"]), (env, name)
            assert np.array_equal(got["lev_dist_vs_true"], want["lev_dist_vs_true"]), (env, name)
        P.check_segment(scorer, oracle, kmers, prob, P.make(*P.SMALL[0], mut=0.5))


def test_startpos_table_geometries(emul_scorer, oracle, kmers, prob, monkeypatch):
    check_startpos_geometries(emul_scorer, oracle, kmers, prob, monkeypatch, n_contigs=1200, L=6000)
