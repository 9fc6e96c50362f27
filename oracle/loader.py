"""TEST INFRASTRUCTURE ONLY: ctypes front-ends of the two CPU checkers.

* ``oracle_calc_breakscore`` -> ``oracle/liboracle.so``: the plain-C restatement
  (``oracle/breakscore_oracle.c``) of upstream ``lib/BreakageScorer.cpp:185-353`` plus the KS
  statistic of ``lib/DeNovoAssembler.R:416-424``.
* ``ref_calc_breakscore`` -> ``oracle/_ref/libref_breakscore*.so``: the UNMODIFIED upstream
  ``calc_breakscore`` compiled from ``/root/reference`` by ``oracle/Makefile``.

Both take the upstream argument list (path, sequencing_reads, true_solution, kmer, bp_kmer,
bp_prob) and return a dict named like the upstream R list (``lib/BreakageScorer.cpp:343-353``).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_c_i64p = C.POINTER(C.c_int64)
_c_i32p = C.POINTER(C.c_int32)
_c_f64p = C.POINTER(C.c_double)


def build(force: bool = False) -> None:
    """Compile liboracle.so (always possible) and oracle/_ref (only where /root/reference exists)."""
    if force or not os.path.exists(os.path.join(_HERE, "liboracle.so")) or (
        os.path.exists("/root/reference/lib/BreakageScorer.cpp")
        and not os.path.exists(os.path.join(_HERE, "_ref", "libref_breakscore.so"))
    ):
        subprocess.run(["make", "-s", "-C", _HERE], check=True)


def flatten(strings):
    """list of str/bytes -> (uint8 chars, int64 offsets[n+1])"""
    bs = [s.encode("ascii") if isinstance(s, str) else bytes(s) for s in strings]
    off = np.zeros(len(bs) + 1, dtype=np.int64)
    if bs:
        np.cumsum([len(b) for b in bs], out=off[1:])
    chars = np.frombuffer(b"".join(bs), dtype=np.uint8).copy() if off[-1] else np.zeros(1, np.uint8)
    return chars, off


def _p(a, typ):
    return None if a is None else a.ctypes.data_as(typ)


_libs = {}


def _load(name):
    if name not in _libs:
        path = os.path.join(_HERE, name)
        if not os.path.exists(path):
            build()
        _libs[name] = C.CDLL(path)
    return _libs[name]


def have_ref() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libref_breakscore.so"))


def _prob_dist_offsets(ct_off, kmer):
    lens = np.diff(ct_off) - kmer + 1
    lens = np.maximum(lens, 0)
    off = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    return off


def oracle_calc_breakscore(path, sequencing_reads, true_solution, kmer, bp_kmer, bp_prob,
                           truth_prob=None, want_pos=False, want_hist=False, want_ks=True,
                           want_lev=False, want_prob_dist=True):
    lib = _load("liboracle.so")
    fn = lib.oracle_calc_breakscore
    fn.restype = C.c_int
    ct, ct_off = flatten(path)
    rd, rd_off = flatten(sequencing_reads)
    tr, _ = flatten([true_solution])
    tlen = len(true_solution)
    km, km_off = flatten(bp_kmer)
    prob = np.ascontiguousarray(bp_prob, dtype=np.float64)
    tprob = None if truth_prob is None else np.ascontiguousarray(truth_prob, dtype=np.float64)
    nc, nr, nt = len(path), len(sequencing_reads), len(bp_kmer)
    pd_off = _prob_dist_offsets(ct_off, kmer)
    out = {
        "sequence_len": np.zeros(nc, np.int32),
        "bp_score": np.zeros(nc, np.float64),
        "bp_score_norm_by_break_freqs": np.zeros(nc, np.float64),
        "bp_score_norm_by_len": np.zeros(nc, np.float64),
        "kmer_breaks": np.zeros(nc, np.int32),
        "path_prob_dist_startpos": np.zeros(nc, np.int32),
    }
    pd = np.zeros(max(int(pd_off[-1]), 1), np.float64) if want_prob_dist else None
    pos = np.zeros((nc, nr), np.int32) if want_pos else None
    hist = np.zeros((nc, nt + 1), np.int32) if want_hist else None
    ksa = np.zeros(nc, np.float64) if want_ks else None
    ksb = np.zeros(nc, np.float64) if want_ks else None
    lev = np.zeros(nc, np.int32) if want_lev else None
    rc = fn(_p(ct, C.c_char_p), _p(ct_off, _c_i64p), C.c_int64(nc),
            _p(rd, C.c_char_p), _p(rd_off, _c_i64p), C.c_int64(nr),
            _p(tr, C.c_char_p), C.c_int64(tlen), C.c_int(kmer),
            _p(km, C.c_char_p), _p(km_off, _c_i64p), _p(prob, _c_f64p), C.c_int64(nt),
            _p(tprob, _c_f64p),
            _p(out["sequence_len"], _c_i32p), _p(out["bp_score"], _c_f64p),
            _p(out["bp_score_norm_by_break_freqs"], _c_f64p), _p(out["bp_score_norm_by_len"], _c_f64p),
            _p(out["kmer_breaks"], _c_i32p), _p(out["path_prob_dist_startpos"], _c_i32p),
            _p(pd, _c_f64p), _p(pd_off, _c_i64p), _p(pos, _c_i32p), _p(hist, _c_i32p),
            _p(ksa, _c_f64p), _p(ksb, _c_f64p), _p(lev, _c_i32p))
    if rc != 0:
        raise RuntimeError(f"oracle_calc_breakscore failed with code {rc}")
    out["sequence"] = list(path)
    if want_prob_dist:
        out["path_prob_dist"] = [pd[pd_off[i]:pd_off[i + 1]].copy() for i in range(nc)]
    if want_pos:
        out["pos"] = pos
    if want_hist:
        out["hist"] = hist
    if want_ks:
        out["ks_stat_prob_dist"] = ksa
        out["ks_stat_path_freq"] = ksb
    if want_lev:
        out["lev_dist_vs_true"] = lev
    return out


def oracle_ks_statistic(x, y) -> float:
    lib = _load("liboracle.so")
    fn = lib.oracle_ks_statistic
    fn.restype = C.c_double
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    return float(fn(_p(x, _c_f64p), C.c_int64(len(x)), _p(y, _c_f64p), C.c_int64(len(y))))


def ref_calc_breakscore(path, sequencing_reads, true_solution, kmer, bp_kmer, bp_prob,
                        edit_distance=False, want_prob_dist=True, lib_path=None):
    """The unmodified upstream calc_breakscore.  ``edit_distance=False`` uses the build whose
    edlib stand-in returns 0 (edit distance is off the scored path and dominates run time)."""
    # lib_path: another build of the same driver (tests/test_rcpp_glue.py wraps the Rcpp glue with it)
    lib = C.CDLL(lib_path) if lib_path else _load(os.path.join("_ref", "libref_breakscore.so" if edit_distance
                                                               else "libref_breakscore_noedit.so"))
    fn = lib.ref_calc_breakscore
    fn.restype = C.c_int
    ct, ct_off = flatten(path)
    rd, rd_off = flatten(sequencing_reads)
    tr, _ = flatten([true_solution])
    km, km_off = flatten(bp_kmer)
    prob = np.ascontiguousarray(bp_prob, dtype=np.float64)
    nc, nr, nt = len(path), len(sequencing_reads), len(bp_kmer)
    pd_off = _prob_dist_offsets(ct_off, kmer)
    out = {
        "sequence_len": np.zeros(nc, np.int32),
        "bp_score": np.zeros(nc, np.float64),
        "bp_score_norm_by_break_freqs": np.zeros(nc, np.float64),
        "bp_score_norm_by_len": np.zeros(nc, np.float64),
        "kmer_breaks": np.zeros(nc, np.int32),
        "lev_dist_vs_true": np.zeros(nc, np.int32),
        "path_prob_dist_startpos": np.zeros(nc, np.int32),
    }
    pd = np.zeros(max(int(pd_off[-1]), 1), np.float64) if want_prob_dist else None
    rc = fn(_p(ct, C.c_char_p), _p(ct_off, _c_i64p), C.c_int64(nc),
            _p(rd, C.c_char_p), _p(rd_off, _c_i64p), C.c_int64(nr),
            _p(tr, C.c_char_p), C.c_int64(len(true_solution)), C.c_int(kmer),
            _p(km, C.c_char_p), _p(km_off, _c_i64p), _p(prob, _c_f64p), C.c_int64(nt),
            _p(out["sequence_len"], _c_i32p), _p(out["bp_score"], _c_f64p),
            _p(out["bp_score_norm_by_break_freqs"], _c_f64p), _p(out["bp_score_norm_by_len"], _c_f64p),
            _p(out["kmer_breaks"], _c_i32p), _p(out["lev_dist_vs_true"], _c_i32p),
            _p(out["path_prob_dist_startpos"], _c_i32p), _p(pd, _c_f64p), _p(pd_off, _c_i64p))
    if rc != 0:
        raise RuntimeError(f"reference calc_breakscore failed with code {rc}")
    out["sequence"] = list(path)
    if want_prob_dist:
        out["path_prob_dist"] = [pd[pd_off[i]:pd_off[i + 1]].copy() for i in range(nc)]
    return out


def ref_assemble_contigs(velvet_contigs, dbg_kmer, seed):
    """The unmodified upstream assemble_contigs (lib/BreakageScorer.cpp:79-174, 20 000 shuffles)."""
    lib = _load(os.path.join("_ref", "libref_breakscore_noedit.so"))
    fn = lib.ref_assemble_contigs
    fn.restype = C.c_int64
    ct, ct_off = flatten(velvet_contigs)
    need = C.c_int64(0)
    n = fn(_p(ct, C.c_char_p), _p(ct_off, _c_i64p), C.c_int64(len(velvet_contigs)), C.c_int(dbg_kmer), C.c_int(seed),
           None, C.c_int64(0), C.byref(need))
    if n < 0:
        raise RuntimeError("reference assemble_contigs threw")
    buf = C.create_string_buffer(max(need.value, 1))
    n = fn(_p(ct, C.c_char_p), _p(ct_off, _c_i64p), C.c_int64(len(velvet_contigs)), C.c_int(dbg_kmer), C.c_int(seed),
           buf, C.c_int64(need.value), C.byref(need))
    items = buf.raw[:need.value].split(b"\n")[:-1] if need.value else []
    assert len(items) == n
    return items
