"""Scaffold sets given as parts of base contigs (bs_score_scaffolds) and the checker shared by the CPU-emulation and
the GPU tests: the compositional path against the SAME library scoring the materialised texts (integers, positions and
histograms bit for bit; fp64 sums to 1e-9, another summation order) and against the oracle."""
import numpy as np

from conftest import assert_same_as_oracle, INT_KEYS, F64_KEYS, RTOL, KS_ATOL
from genomeassembler_dev_b200 import breakscore as B
from genomeassembler_dev_b200 import synth, tables

FULL = B.DEFAULT_FLAGS | B.WANT_HIST | B.WANT_POS


def mode_flags(mode):
    """without path_prob_dist out, the KS statistic of the window probabilities comes from the parts too (k_ks_compose)"""
    return FULL & ~B.WANT_PROB_DIST if mode.startswith("ks_from_parts") else FULL


def make_set(seed, length=3000, read_len=40, coverage=10, n_base=8, n_scaffolds=40, overlap=9, max_parts=None, ragged=False,
             mutate=0.0, lib_path=None):
    """truth, reads and a scaffold set: base contigs = consecutive truth intervals, each extended by `overlap` bases of
    its successor (so i -> i+1 is a true suffix/prefix overlap); scaffolds = random chains, the overlap used where the
    chain steps from i to i+1 (sometimes), plain concatenation elsewhere."""
    rng = np.random.default_rng(seed)
    prob8 = tables.sub_table(tables.normalised(tables.load_raw()), 8)
    codes = synth.random_truth_codes(rng, length)
    truth = synth.codes_to_ascii(codes)
    starts = synth.sample_read_starts(rng, codes, read_len, coverage, prob8)
    if ragged:
        lens = rng.integers(max(2, read_len // 3), read_len + 1, size=len(starts))
        reads = [truth[s:s + l].tobytes() for s, l in zip(starts, lens)]
    else:
        reads = [truth[s:s + read_len].tobytes() for s in starts]
    cuts = np.sort(rng.choice(np.arange(1, length // 16), size=n_base - 1, replace=False)) * 16
    bounds = np.concatenate([[0], cuts, [length]])
    base = []
    for a, b in zip(bounds[:-1], bounds[1:]):
        c = truth[a:min(b + overlap, length)].copy()
        if mutate and rng.random() < mutate and len(c) > 20:
            p = int(rng.integers(overlap + 1, len(c) - overlap - 1))
            c[p] = ord("ACGT"[("ACGT".index(chr(c[p])) + 1) % 4])
        base.append(c.tobytes())
    part_start, part_base, part_ov = [0], [], []
    for _ in range(n_scaffolds):
        k = int(rng.integers(1, (max_parts or n_base) + 1))
        chain = [int(rng.integers(0, n_base))]
        while len(chain) < k:
            if chain[-1] + 1 < n_base and rng.random() < 0.5:
                chain.append(chain[-1] + 1)
            else:
                chain.append(int(rng.integers(0, n_base)))
        for i, bidx in enumerate(chain):
            ov = 0
            if i > 0 and bidx == chain[i - 1] + 1 and overlap < len(base[bidx]) and rng.random() < 0.8:
                ov = overlap
            part_base.append(bidx)
            part_ov.append(ov)
        part_start.append(len(part_base))
    sset = B.ScaffoldSet(base, part_start, part_base, part_ov, lib_path)
    return truth.tobytes(), reads, sset


def check_scaffolds(scorer, oracle, kmers, prob, truth, reads, sset, kmer=8, flags=FULL, oracle_sample=None, second=None):
    """compositional == rescan of the texts (same library) == oracle (optionally on a sample of the scaffolds)"""
    scorer.set_table(kmers, prob)
    if second is not None:
        scorer.set_second_table(second)
        flags |= B.WANT_SECOND_TABLE
    texts = sset.texts()
    got = scorer.score_scaffolds(sset, reads, truth, kmer=kmer, flags=flags)
    assert got["sequence"] == texts
    ref = scorer.score(texts, reads, truth, kmer=kmer, flags=flags)
    keys_i = list(INT_KEYS) + (["pos"] if flags & B.WANT_POS else []) + (["hist"] if flags & B.WANT_HIST else [])
    for k in keys_i:
        assert np.array_equal(got[k], ref[k]), (k, np.flatnonzero(np.any(np.atleast_2d(np.asarray(got[k]).T != np.asarray(ref[k]).T), axis=0))[:10])
    f64 = list(F64_KEYS) + (["bp_score2", "bp_score_norm_by_break_freqs2", "bp_score_norm_by_len2"] if second is not None else [])
    for k in f64:
        np.testing.assert_allclose(got[k], ref[k], rtol=RTOL, atol=0, equal_nan=True, err_msg=k)
    if flags & B.WANT_KS:
        ks = ["ks_stat_prob_dist", "ks_stat_path_freq"] + (["ks_stat_prob_dist2", "ks_stat_path_freq2"] if second is not None else [])
        for k in ks:
            np.testing.assert_allclose(got[k], ref[k], rtol=RTOL, atol=KS_ATOL, equal_nan=True, err_msg=k)
        assert np.array_equal(got["ks_stat_prob_dist"], ref["ks_stat_prob_dist"], equal_nan=True)  # (integer numerators: exact)
    if flags & B.WANT_PROB_DIST:
        for a, b in zip(got["path_prob_dist"], ref["path_prob_dist"]):
            assert np.array_equal(a, b)
    if oracle is not None:
        idx = list(range(len(texts))) if oracle_sample is None else list(oracle_sample)
        want = oracle.oracle_calc_breakscore([texts[i] for i in idx], reads, truth, kmer, kmers, prob,
                                             want_pos=bool(flags & B.WANT_POS), want_hist=bool(flags & B.WANT_HIST))
        sub = {}
        for k, v in got.items():
            if k in ("sequence", "path_prob_dist", "path_prob_dist2"):
                sub[k] = [v[i] for i in idx]
            elif isinstance(v, np.ndarray) and len(v) == len(texts):
                sub[k] = v[idx]
        assert_same_as_oracle(sub, want, check_pos=bool(flags & B.WANT_POS), check_hist=bool(flags & B.WANT_HIST))
    return got


def hand_sets():
    """(name, base contigs, chains [(base, overlap), ...], reads, truth, kmer): degenerate and adversarial shapes"""
    t = b"ACGTTGCAAGGCTTACCGATAGGATCCGATTACAGGCATTAGCCGATAGACCATTGGCAAGTCCGATAGGCTAAGCTTGGCAATCGGATACCAGT"
    a, b, c, d = t[0:30], t[22:60], t[52:80], t[80:96]   # a|b overlap 8, b|c overlap 8, c|d overlap 0
    rep = b"ACGGTCA" * 12
    out = []
    reads1 = [t[i:i + 12] for i in range(0, 84, 3)] + [t[20:40], t[50:62], t[76:90], t[25:28], b"", t[0:30]]
    out.append(("chain_with_overlaps", [a, b, c, d], [[(0, 0), (1, 8), (2, 8), (3, 0)], [(1, 0), (2, 8)], [(3, 0), (0, 0)], [(2, 0)]],
                reads1, t, 8))
    # parts shorter than the reads: one read crosses several junctions
    s = [t[0:5], t[5:9], t[9:16], t[16:40], t[40:43], t[43:96]]
    out.append(("parts_shorter_than_reads", s, [[(0, 0), (1, 0), (2, 0), (3, 0), (4, 0), (5, 0)], [(4, 0), (0, 0), (1, 0), (2, 0)],
                                                 [(1, 0), (2, 0), (3, 0)]],
                [t[i:i + 14] for i in range(0, 80, 2)] + [t[2:30], t[3:8], t[4:6]], t, 8))
    # the same base contig twice, and repeats inside: the leftmost rule across parts
    out.append(("repeated_parts", [rep[:40], a, rep[7:35]], [[(0, 0), (1, 0), (0, 0)], [(1, 0), (2, 0), (0, 0), (1, 0)], [(2, 0), (2, 0)]],
                [rep[i:i + 10] for i in range(0, 30)] + [a[20:30] + rep[:6], rep[30:40] + a[:5], a[5:17]], t + rep, 8))
    # bytes outside ACGT in base contigs and reads (text comparison paths), reads crossing junctions there
    n1, n2 = b"ACGNTGCAAGGCTTNNCGATAGGA", b"TTGACCNAGGTCAGGTACCA"
    out.append(("non_acgt", [n1, n2, a], [[(0, 0), (1, 0), (2, 0)], [(1, 0), (0, 0)], [(2, 0), (1, 0), (0, 0)]],
                [b"GNTGC", b"TNNCG", b"NN", b"AGGATTGA", b"GGATTGACCNAG", b"CNAGG", b"CCAACGNTG", b"TACCAACGT", n1[-6:] + n2[:7],
                 b"acgt", a[-5:] + n2[:4], b"N", b"GTACCAACGNT"], t, 8))
    out.append(("no_reads", [a, b], [[(0, 0), (1, 8)], [(1, 0)]], [], t, 8))
    out.append(("single_parts_other_kmer", [a, b, c], [[(0, 0)], [(1, 0)], [(2, 0)], [(0, 0), (1, 8), (2, 8)]],
                [t[i:i + 9] for i in range(0, 70, 2)], t, 4))
    out.append(("kmer_longer_than_rows", [a, b], [[(0, 0), (1, 8)], [(1, 0), (0, 0)]], [t[i:i + 9] for i in range(0, 50, 2)], t, 10))
    # ragged reads, the shortest one shorter than a word: seed length < 32, junction windows as wide as the longest read
    out.append(("ragged_reads", [t[0:40], t[32:96]], [[(0, 0), (1, 8)], [(1, 0), (0, 0)]],
                [t[i:i + 3 + (i % 37)] for i in range(0, 60)] + [t[1:70]], t, 8))
    # one break k-mer more than CC_DENSE (1024) times in a scaffold: the running tallies of KS-B give way to the sweep
    out.append(("one_kmer_a_thousand_times", [b"A" * 300, b"ACGT" * 10 + b"AAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA"],
                [[(0, 0), (1, 0)], [(1, 0), (0, 0)], [(1, 0)]], [b"A" * 20] * 1100 + [b"AAAC", b"CGTA", b"GTAAAAAA"], b"A" * 300 + b"ACGT" * 10, 8))
    # a read of ACGT only between reads with Ns (the packers flag it as well: "next to a byte outside ACGT"), crossing several
    # junctions of a set that is kept as packed words alone
    out.append(("flagged_pure_read_crossing_short_parts", [b"CTGGGCG"], [[(0, 0)], [(0, 0)] * 5, [(0, 0)] * 4],
                [b"AATTNTGGTGCGCCNATANANNGNCGAGC", b"GCGCTGGGCGCTGGGCGCTGGGCGCTGGG", b"NCTNTTACATGCCGGCGTTTGNTTNAANN"],
                b"GANNACGATTATCTGGGCGTGANNNT", 8))
    return out


def hand_scaffold_set(base, chains, lib_path=None):
    part_start, part_base, part_ov = [0], [], []
    for ch in chains:
        for bidx, ov in ch:
            part_base.append(bidx)
            part_ov.append(ov)
        part_start.append(len(part_base))
    return B.ScaffoldSet(base, part_start, part_base, part_ov, lib_path)


def check_set_independence(scorer, kmers, prob, monkeypatch, lib_path=None, **kw):
    """a scaffold's record is the same bits whatever else travels in the call (any subset of the set: what a rank of a
    sharded job scores) and whether or not its junctions have lists"""
    from genomeassembler_dev_b200 import sharding
    truth, reads, sset = make_set(lib_path=lib_path, **kw)
    scorer.set_table(kmers, prob)
    flags = B.WANT_KS | B.WANT_STARTPOS
    whole = scorer.score_scaffolds(sset, reads, truth, flags=flags)
    keys = sharding.RECORD_F64 + sharding.RECORD_I32
    for world in (2, 3, 5):
        for part in sharding.shard_contigs_lpt(sset.lengths(), world):
            if len(part):
                loc = scorer.score_scaffolds(sset.subset(part), reads, truth, flags=flags)
                for k in keys:
                    assert np.array_equal(loc[k], whole[k][part], equal_nan=True), (world, k)
    monkeypatch.setenv("BS_COMPOSE_JUNCTIONS", "0")
    probed = scorer.score_scaffolds(sset, reads, truth, flags=flags)
    for k in keys:
        assert np.array_equal(probed[k], whole[k], equal_nan=True), ("junctions probed", k)
