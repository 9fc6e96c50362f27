"""Breakage-probability tables: the ``bp_kmer`` / ``bp_prob`` arguments of ``calc_breakscore``.

The upstream R driver builds them in ``lib/GenerateReads.R:153-184``: the four CSVs
``data/QueryTable/QueryTable_kmer-{2,4,6,8}.csv`` are concatenated in that order and the
probabilities are normalised JOINTLY over all 69 904 rows (``prob / sum(prob)``).  The
"random" pass of ``lib/DeNovoAssembler.R:326-333`` replaces them by ``1/69904`` everywhere.
"""
from __future__ import annotations

import itertools
import os

import numpy as np

KS = (2, 4, 6, 8)
N_ROWS = sum(4 ** k for k in KS)  # 69 904
_FIXTURE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                        "tests", "golden", "query_table_k2468.npz")


def kmer_strings(k: int):
    """All 4^k ACGT strings in lexicographic order (the row order of the upstream CSVs)."""
    return ["".join(t) for t in itertools.product("ACGT", repeat=k)]


def all_kmer_strings():
    out = []
    for k in KS:
        out.extend(kmer_strings(k))
    return out


def load_raw_from_csv(table_dir: str) -> np.ndarray:
    """Raw (un-normalised) probabilities in row order 2-,4-,6-,8-mers, checked to be lexicographic."""
    vals = []
    for k in KS:
        names, probs = [], []
        with open(os.path.join(table_dir, f"QueryTable_kmer-{k}.csv")) as fh:
            header = fh.readline().strip()
            assert header == "kmer,prob", header
            for line in fh:
                a, b = line.strip().split(",")
                names.append(a)
                probs.append(float(b))
        assert names == kmer_strings(k), f"k={k}: rows are not in lexicographic ACGT order"
        vals.append(np.asarray(probs, dtype=np.float64))
    return np.concatenate(vals)


def load_raw(path: str | None = None) -> np.ndarray:
    """Raw probabilities from the committed fixture (tests/golden/make_table_fixture.py made it)."""
    with np.load(path or _FIXTURE) as z:
        return z["raw_prob"].astype(np.float64)


def normalised(raw: np.ndarray) -> np.ndarray:
    """lib/GenerateReads.R:173-176: divide by the grand total over all four tables."""
    return raw / raw.sum()


def uniform(n: int = N_ROWS) -> np.ndarray:
    """lib/DeNovoAssembler.R:326-330: rep(1/length, length)."""
    return np.full(n, 1.0 / n, dtype=np.float64)


def sub_table(prob_all: np.ndarray, k: int) -> np.ndarray:
    """Rows of the joint table that belong to k-mers of length k (df_prob$kmer_<k>)."""
    start = 0
    for kk in KS:
        if kk == k:
            return prob_all[start:start + 4 ** kk]
        start += 4 ** kk
    raise ValueError(f"no rows for k={k}")
