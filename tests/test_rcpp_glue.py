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
                    os.path.join(ROOT, "oracle", "ref_driver.cpp"), "-o", out,
                    "-L", pkg, "-lbreakscore", f"-Wl,-rpath,{pkg}"], check=True)
    return out


def test_glue_compiles_and_links(glue_lib):
    syms = subprocess.run(["nm", "-D", "--defined-only", glue_lib], capture_output=True, text=True, check=True).stdout
    assert "ref_calc_breakscore" in syms  # the driver's wrapper around the glue's calc_breakscore
    assert "ref_assemble_contigs" in syms  # ... and around its assemble_contigs


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

