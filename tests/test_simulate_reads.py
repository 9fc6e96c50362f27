"""Read simulation on the device (upstream lib/GenerateReads.R:302-313,368-379).  The random stream
is the library's own, so the checks are about the LAW: every read is a substring of its truth, the
number of draws and the drop rule are upstream's, the start positions follow the table probabilities
(chi-square against the exact expectation), and the stream is reproducible.  CPU: kernel sources
under the emulation; GPU (-m gpu): the nvcc build at the cfg-2 segment size."""
import numpy as np
import pytest
from scipy import stats

from genomeassembler_dev_b200 import synth, tables


def check_law(scorer, kmers, prob, L, r, cov, seed, n_seg):
    scorer.set_table(kmers, prob)
    rng = np.random.default_rng(seed)
    codes = [synth.random_truth_codes(rng, L) for _ in range(n_seg)]
    truths = [synth.codes_to_ascii(c).tobytes() for c in codes]
    reads, srs = scorer.simulate_reads(truths, r, cov, seed=seed)
    again, srs2 = scorer.simulate_reads(truths, r, cov, seed=seed)
    assert np.array_equal(reads, again) and np.array_equal(srs, srs2)          # reproducible
    other, _ = scorer.simulate_reads(truths, r, cov, seed=seed + 1)
    assert other.shape != reads.shape or not np.array_equal(other, reads)       # seeded
    p8 = tables.sub_table(prob, 8)
    n_draw = int(np.ceil(cov * L / r))
    for s in range(n_seg):
        seg_reads = reads[srs[s]:srs[s + 1]]
        w = p8[synth.rolling_codes(codes[s], 8)]
        w = w / w.sum()
        keep = np.arange(len(w)) + r <= L                                       # lib/GenerateReads.R:310-313
        # count: binomial(n_draw, P(keep))
        pk = w[keep].sum()
        assert abs(len(seg_reads) - n_draw * pk) <= 6 * np.sqrt(n_draw * pk * (1 - pk)) + 1
        # every read is the substring at its start; starts recovered by search (32+ random bases are unique)
        t = truths[s]
        starts = np.array([t.find(x.tobytes()) for x in seg_reads])
        assert (starts >= 0).all() and (starts + r <= L).all()
        if r >= 24:
            # start distribution: chi-square over 40 equal-mass bins of the kept windows
            cw = np.cumsum(np.where(keep, w, 0.0))
            edges = np.searchsorted(cw, np.linspace(0, cw[-1], 41)[1:-1])
            obs = np.bincount(np.searchsorted(edges, starts, side="right"), minlength=40)
            exp = np.diff(np.concatenate([[0.0], cw[edges], [cw[-1]]])) / cw[-1] * len(starts)
            chi2 = ((obs - exp) ** 2 / exp).sum()
            assert stats.chi2.sf(chi2, 39) > 1e-5, (chi2, obs, exp)


def test_simulated_reads_follow_upstream_law_emulated(emul_scorer, kmers, prob):
    check_law(emul_scorer, kmers, prob, L=3000, r=40, cov=12, seed=5, n_seg=2)


def test_edge_shapes_emulated(emul_scorer, kmers, prob):
    emul_scorer.set_table(kmers, prob)
    reads, srs = emul_scorer.simulate_reads([b"ACGTACGTAC", b"", b"ACGTTGCAAGGCTTACCGATAGGA"], 12, 5, seed=1)
    assert srs[1] == 0 and srs[2] == 0            # 10 < 12: every draw overruns; empty truth: no draws
    assert reads.shape[1] == 12 and all(x.tobytes() in b"ACGTTGCAAGGCTTACCGATAGGA" for x in reads)
    reads, srs = emul_scorer.simulate_reads([b"NNNNNNNNNNNNNNNN"], 4, 5, seed=1)
    assert len(reads) == 0                         # no window is a table row: nothing to draw from


@pytest.mark.gpu
def test_simulated_reads_follow_upstream_law_gpu(gpu_scorer, kmers, prob):
    check_law(gpu_scorer, kmers, prob, L=50000, r=150, cov=30, seed=1234, n_seg=3)
    check_law(gpu_scorer, kmers, prob, L=20000, r=40, cov=40, seed=7, n_seg=2)


@pytest.mark.gpu
def test_simulate_then_score_on_device_inputs(gpu_scorer, oracle, kmers, prob):
    """the simulated reads are valid scorer input: score them against exact contigs, compare with the oracle"""
    import parity_cases as P
    gpu_scorer.set_table(kmers, prob)
    seg = synth.make_segment(55, length=20000, read_len=100, coverage=5, n_contigs=8, mut_frac=0.2)
    reads, srs = gpu_scorer.simulate_reads([seg.truth], 100, 20, seed=99)
    P.check_segment(gpu_scorer, oracle, kmers, prob, synth.Segment(seg.truth, reads, seg.contigs))
