"""The Rcpp front-end (rcpp/BreakageScorer.cpp) cannot be built against real Rcpp here (no R in
the image), so it is compiled against the oracle's stand-in Rcpp.h and driven through the same
C wrapper as the verbatim reference build (oracle/ref_driver.cpp).  CPU: it compiles and links
against libbreakscore.so.  GPU: its calc_breakscore returns the reference's list members on the
reference vectors."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, assert_matches_reference


@pytest.fixture(scope="module")
def glue_lib(product_lib, tmp_path_factory):
    out = str(tmp_path_factory.mktemp("glue") / "libglue_breakscore.so")
    pkg = os.path.dirname(product_lib)
    subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-w",
                    "-I", os.path.join(ROOT, "oracle", "shim", "inc"), "-I", os.path.join(ROOT, "include"),
                    f'-DBS_REFERENCE_SRC="{os.path.join(ROOT, "rcpp", "BreakageScorer.cpp")}"',
                    os.path.join(ROOT, "oracle", "ref_driver.cpp"), os.path.join(ROOT, "tests", "glue_scaffold_driver.cpp"), "-o", out,
                    "-L", pkg, "-lbreakscore", f"-Wl,-rpath,{pkg}"], check=True)
    return out


def test_glue_compiles_and_links(glue_lib):
    syms = subprocess.run(["nm", "-D", "--defined-only", glue_lib], capture_output=True, text=True, check=True).stdout
    assert "ref_calc_breakscore" in syms  # the driver's wrapper around the glue's calc_breakscore
    assert "ref_assemble_contigs" in syms  # ... and around its assemble_contigs
    assert "glue_assemble_and_score" in syms  # ... and (tests/glue_scaffold_driver.cpp) around its assemble_and_score


def test_glue_assemble_contigs_equals_reference_vectors(glue_lib):
    """the glue's assemble_contigs export (host code, no GPU) on the upstream-generated vectors"""
    import ctypes as C
    from test_assemble_contigs import CASES
    from oracle import loader as O
    lib = C.CDLL(glue_lib)
    fn = lib.ref_assemble_contigs
    fn.restype = C.c_int64
    for name, contigs, k, seed, expected, threw in CASES:
        ct, off = O.flatten(contigs)
        need = C.c_int64(0)
        n = fn(ct.ctypes.data_as(C.c_char_p), off.ctypes.data_as(C.POINTER(C.c_int64)), C.c_int64(len(contigs)),
               C.c_int(k), C.c_int(seed), None, C.c_int64(0), C.byref(need))
        if threw:
            assert n == -1, name
            continue
        buf = C.create_string_buffer(max(need.value, 1))
        n = fn(ct.ctypes.data_as(C.c_char_p), off.ctypes.data_as(C.POINTER(C.c_int64)), C.c_int64(len(contigs)),
               C.c_int(k), C.c_int(seed), buf, C.c_int64(need.value), C.byref(need))
        got = buf.raw[:need.value].split(b"\n")[:-1] if need.value else []
        assert got == expected, name
    assert "libbreakscore.so" in subprocess.run(["ldd", glue_lib], capture_output=True, text=True).stdout


@pytest.mark.gpu
def test_glue_returns_the_reference_list(glue_lib, oracle, kmers, table_set, ref_vectors):
    for case in ref_vectors:
        if case["table"] not in ("real", "rowid"):
            continue
        got = oracle.ref_calc_breakscore(case["path"], case["reads"], case["truth"], case["kmer"], kmers,
                                         table_set[case["table"]], lib_path=glue_lib)
        assert_matches_reference(got, case["expected"])
        # edit distance: the glue asks for it (upstream returns edlib's HW distance); checked against the oracle's DP
        want = oracle.oracle_calc_breakscore(case["path"], case["reads"], case["truth"], case["kmer"], kmers,
                                             table_set[case["table"]], want_ks=False, want_lev=True, want_prob_dist=False)
        assert np.array_equal(got["lev_dist_vs_true"], want["lev_dist_vs_true"]), case["name"]



@pytest.mark.gpu
def test_glue_assemble_and_score(glue_lib, gpu_scorer, oracle, kmers, prob):
    """assemble_and_score == assemble_contigs followed by calc_breakscore (lib/DeNovoAssembler.R:343-355): the strings of
    the scaffold explosion in upstream's order, and for them the list members of the scorer (here against the Python
    binding scoring the same texts the ordinary way, itself checked against the oracle elsewhere)."""
    import ctypes as C
    from genomeassembler_dev_b200 import breakscore as B, synth
    from oracle import loader as O
    seg = synth.make_segment(78, length=6000, read_len=60, coverage=15, n_contigs=6, mut_frac=0.0)
    contigs = [seg.truth[s:s + len(c) + 50] for s, c in zip(seg.contig_truth_start, seg.contigs)]
    lib = C.CDLL(glue_lib)
    fn = lib.glue_assemble_and_score
    fn.restype = C.c_int64
    ct, ct_off = O.flatten(contigs)
    rd, rd_off = O.flatten(seg.read_list)
    km, km_off = O.flatten(kmers)
    pr = np.ascontiguousarray(prob, np.float64)
    p = lambda a, t: a.ctypes.data_as(t)  # noqa: E731
    i64p, i32p, f64p = C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_double)
    head = [p(ct, C.c_char_p), p(ct_off, i64p), C.c_int64(len(contigs)), C.c_int(13), C.c_int(1234), p(rd, C.c_char_p), p(rd_off, i64p),
            C.c_int64(len(seg.read_list)), C.c_char_p(seg.truth), C.c_int64(len(seg.truth)), C.c_int(8), p(km, C.c_char_p), p(km_off, i64p),
            p(pr, f64p), C.c_int64(len(kmers))]
    nbytes = C.c_int64(0)
    n = fn(*head, C.c_int64(0), None, None, None, None, None, None, None, None, C.c_int64(0), C.byref(nbytes))
    assert n > 0
    i32 = [np.zeros(n, np.int32) for _ in range(4)]
    f64 = [np.zeros(n, np.float64) for _ in range(3)]
    buf = C.create_string_buffer(nbytes.value)
    assert fn(*head, C.c_int64(n), *[p(a, i32p) for a in i32], *[p(a, f64p) for a in f64], buf, C.c_int64(nbytes.value), C.byref(nbytes)) == n
    strings = buf.raw[:nbytes.value].split(b"\n")[:-1]
    assert strings == B.assemble_contigs(contigs, 13, 1234)
    gpu_scorer.set_table(kmers, prob)
    want = gpu_scorer.score(strings, seg.read_list, seg.truth, flags=B.DEFAULT_FLAGS | B.WANT_LEV)
    for got, key in zip(i32, ("sequence_len", "kmer_breaks", "path_prob_dist_startpos", "lev_dist_vs_true")):
        assert np.array_equal(got, want[key]), key
    np.testing.assert_allclose(f64[0], want["bp_score"], rtol=1e-9, atol=0)
    assert np.array_equal(f64[1], want["ks_stat_prob_dist"], equal_nan=True)
    np.testing.assert_allclose(f64[2], want["ks_stat_path_freq"], rtol=1e-9, atol=1e-12, equal_nan=True)
