"""Pins the oracle (oracle/breakscore_oracle.c) to outputs of the UNMODIFIED upstream
calc_breakscore (tests/golden/ref_vectors.npz, made by tests/golden/make_ref_vectors.py) and to the
known-answer vectors of SURVEY.md appendix A.4."""
import numpy as np
import pytest

from conftest import assert_matches_reference, load_ref_vectors

CASES = load_ref_vectors()


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_matches_reference_vector(case, oracle, kmers, table_set):
    got = oracle.oracle_calc_breakscore(case["path"], case["reads"], case["truth"], case["kmer"], kmers,
                                        table_set[case["table"]], want_ks=False)
    assert_matches_reference(got, case["expected"])
    if case["table"] in ("rowid", "rowsq"):
        # integer tables: bp_score is an exact integer checksum of the break histogram
        assert np.array_equal(got["bp_score"], case["expected"]["bp_score"])


# which table row a read placed at `pos` increments
KAT_BINS = {
    "kat1_pos0": "ACGTTGCA", "kat1_pos1": "AC", "kat1_pos2": "ACGT", "kat1_pos3": "ACGTTG",
    "kat1_pos4": "ACGTTGCA", "kat1_pos5": "CGTTGCAA", "kat1_pos6": "GTTGCAAG",
    "kat2_k4_pos0": "ACGTTGCA", "kat2_k4_pos1": "AC", "kat2_k4_pos2": "ACGT", "kat2_k4_pos3": "CGTTGCAA",
    "kat2_k4_pos5": "TTGCAAGG", "kat3_leftmost": "TTTTACGT",
}


@pytest.mark.parametrize("name", sorted(KAT_BINS))
def test_known_answer_bins(name, oracle, kmers, table_set, ref_vectors):
    case = next(c for c in ref_vectors if c["name"] == f"{name}/rowid")
    row = kmers.index(KAT_BINS[name])
    mult = len(case["reads"])
    # the reference itself: bp_score = (row + 1) * multiplicity under the rowid table
    assert case["expected"]["bp_score"][0] == (row + 1) * mult
    assert case["expected"]["kmer_breaks"][0] == mult
    got = oracle.oracle_calc_breakscore(case["path"], case["reads"], case["truth"], case["kmer"], kmers,
                                        table_set["rowid"], want_hist=True, want_ks=False)
    hist = got["hist"][0]
    assert hist[row] == mult and hist.sum() == mult


def test_known_answer_n_rule(oracle, kmers, table_set, ref_vectors):
    """A.4-4: a break window holding N is not in the table: counted in kmer_breaks, score 0."""
    case = next(c for c in ref_vectors if c["name"] == "kat4_N/rowid")
    assert case["expected"]["kmer_breaks"][0] == 1 and case["expected"]["bp_score"][0] == 0.0
    got = oracle.oracle_calc_breakscore(case["path"], case["reads"], case["truth"], 8, kmers, table_set["rowid"],
                                        want_hist=True, want_ks=False)
    assert got["hist"][0][-1] == 1 and got["hist"][0].sum() == 1


def test_known_answer_startpos(ref_vectors):
    """A.4-5: truth.find(contig) only when a read hit, else 0."""
    by = {c["name"]: c["expected"] for c in ref_vectors}
    assert by["kat5_startpos4/real"]["path_prob_dist_startpos"][0] == 4
    assert by["kat5_startpos_absent/real"]["path_prob_dist_startpos"][0] == -1
    assert by["kat5_nohit/real"]["path_prob_dist_startpos"][0] == 0
    assert by["kat5_nohit/real"]["kmer_breaks"][0] == 0
