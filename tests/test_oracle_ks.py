"""The KS statistic of the oracle (restating R's stats::ks.test, lib/DeNovoAssembler.R:416-424;
PARITY UNPINNED by upstream tests, R absent) cross-checked against scipy.stats.ks_2samp."""
import numpy as np
import pytest
from scipy import stats


@pytest.mark.parametrize("seed", range(8))
def test_ks_vs_scipy(seed, oracle):
    rng = np.random.default_rng(seed)
    nx, ny = int(rng.integers(1, 400)), int(rng.integers(1, 400))
    # heavy ties, like table probabilities (reverse-complement pairs share values)
    pool = rng.random(50)
    x = rng.choice(pool, nx) if seed % 2 else rng.random(nx)
    y = rng.choice(pool, ny)
    d = oracle.oracle_ks_statistic(x, y)
    assert d == pytest.approx(stats.ks_2samp(x, y, method="asymp").statistic, rel=1e-12, abs=1e-15)


def test_ks_drops_nan_and_handles_empty(oracle):
    x = np.array([0.1, np.nan, 0.3]); y = np.array([np.nan, 0.2, 0.2, 0.5])
    assert oracle.oracle_ks_statistic(x, y) == pytest.approx(stats.ks_2samp([0.1, 0.3], [0.2, 0.2, 0.5]).statistic)
    assert np.isnan(oracle.oracle_ks_statistic(np.array([np.nan]), y))


def test_ks_columns_of_the_scorer(oracle, kmers, prob):
    """ks_stat_prob_dist == KS(path_prob_dist, truth windows); ks_stat_path_freq == KS(hist/total, same)."""
    from genomeassembler_dev_b200 import synth, tables
    seg = synth.make_segment(3, length=3000, read_len=50, coverage=10, n_contigs=5, n_gap_scaffolds=1)
    res = oracle.oracle_calc_breakscore(seg.contigs, seg.read_list, seg.truth, 8, kmers, prob, want_hist=True)
    p8 = tables.sub_table(prob, 8)
    codes = np.searchsorted(np.frombuffer(b"ACGT", np.uint8), np.frombuffer(seg.truth, np.uint8))
    y = p8[synth.rolling_codes(codes.astype(np.uint8), 8)]
    for c in range(len(seg.contigs)):
        a = stats.ks_2samp(res["path_prob_dist"][c], y, method="asymp").statistic
        assert res["ks_stat_prob_dist"][c] == pytest.approx(a, rel=1e-12)
        tot = res["kmer_breaks"][c]
        if tot:
            b = stats.ks_2samp(res["hist"][c][:-1] / tot, y, method="asymp").statistic
            assert res["ks_stat_path_freq"][c] == pytest.approx(b, rel=1e-12)
