"""Live differential test: the oracle restatement vs the UNMODIFIED upstream scorer compiled into
oracle/_ref (present where /root/reference was available at build time; the .so travels to the
GPU box).  Skipped when oracle/_ref is absent."""
import numpy as np
import pytest

from conftest import assert_matches_reference
from genomeassembler_dev_b200 import synth, tables


@pytest.mark.parametrize("seed,L,r,cov,nc,gaps", [
    (1, 5000, 100, 10, 8, 1), (2, 6000, 150, 6, 6, 0), (3, 2000, 14, 20, 10, 1), (4, 3000, 40, 20, 5, 2),
    (5, 8000, 64, 5, 12, 0), (6, 1200, 300, 10, 3, 0),
])
def test_oracle_equals_reference_build(seed, L, r, cov, nc, gaps, oracle, kmers, table_set):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (no upstream checkout at build time)")
    seg = synth.make_segment(seed, length=L, read_len=r, coverage=cov, n_contigs=nc,
                             prob8=tables.sub_table(table_set["real"], 8), mut_frac=0.3, n_gap_scaffolds=gaps)
    for tab in ("real", "rowid"):
        ref = oracle.ref_calc_breakscore(seg.contigs, seg.read_list, seg.truth, 8, kmers, table_set[tab])
        got = oracle.oracle_calc_breakscore(seg.contigs, seg.read_list, seg.truth, 8, kmers, table_set[tab], want_ks=False)
        expected = dict(ref)
        expected["path_prob_dist"] = np.concatenate(ref["path_prob_dist"])
        assert_matches_reference(got, expected)
        if tab == "rowid":
            assert np.array_equal(got["bp_score"], ref["bp_score"])


def test_variable_length_reads(oracle, kmers, table_set):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built")
    seg = synth.make_segment(11, length=3000, read_len=60, coverage=8, n_contigs=5, mut_frac=0.2)
    rng = np.random.default_rng(0)
    reads = [r[: int(rng.integers(9, 61))] for r in seg.read_list] + [b"", b"ACGTN", b"NNNN"]
    ref = oracle.ref_calc_breakscore(seg.contigs, reads, seg.truth, 8, kmers, table_set["rowid"])
    got = oracle.oracle_calc_breakscore(seg.contigs, reads, seg.truth, 8, kmers, table_set["rowid"], want_ks=False)
    expected = dict(ref)
    expected["path_prob_dist"] = np.concatenate(ref["path_prob_dist"])
    assert_matches_reference(got, expected)
