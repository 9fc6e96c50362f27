"""Seeded synthetic workloads of the shapes named in BASELINE.json.

Read simulation mirrors the upstream R code (``lib/GenerateReads.R:243-259,302-313,368-379``):
start positions are sampled WITH replacement from the ``L-7`` rolling-octamer positions with
probability proportional to the table probability of the octamer starting there,
``ceil(coverage * L / read_len)`` draws, and draws whose read would overrun the segment are
dropped (so every read has exactly ``read_len`` bases).  Contig sets are "velvet-style": exact
substrings of the segment, a fraction carrying one substitution (so they are not substrings
any more: ``path_prob_dist_startpos == -1``) and optionally an ``N``-gap scaffold
(``velvetg -scaffolding yes``, ``lib/DeNovoAssembler.R:217``).

Everything here is host-side input generation for tests and benchmarks; no scoring happens.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import tables

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def random_truth_codes(rng: np.random.Generator, length: int) -> np.ndarray:
    """i.i.d. uniform 2-bit codes (0..3 = A,C,G,T)."""
    return rng.integers(0, 4, size=length, dtype=np.uint8)


def codes_to_ascii(codes: np.ndarray) -> np.ndarray:
    return _ACGT[codes]


def rolling_codes(codes: np.ndarray, k: int) -> np.ndarray:
    """Lexicographic index of every k-window (length len-k+1, int64)."""
    n = len(codes) - k + 1
    if n <= 0:
        return np.zeros(0, dtype=np.int64)
    out = np.zeros(n, dtype=np.int64)
    for i in range(k):
        out = out * 4 + codes[i:i + n]
    return out


def sample_read_starts(rng, truth_codes, read_len, coverage, prob8):
    """lib/GenerateReads.R:302-313 (0-based starts)."""
    length = len(truth_codes)
    w = prob8[rolling_codes(truth_codes, 8)]
    n_draw = int(np.ceil(coverage * length / read_len))
    starts = rng.choice(len(w), size=n_draw, replace=True, p=w / w.sum())
    return starts[starts + read_len <= length]


def velvet_style_intervals(rng, length, n_contigs, min_len, gap_max=40):
    """Cut [0, length) into n_contigs pieces, trim random gaps, drop pieces below min_len."""
    cuts = np.sort(rng.choice(np.arange(1, length), size=max(n_contigs - 1, 0), replace=False))
    bounds = np.concatenate([[0], cuts, [length]])
    out = []
    for a, b in zip(bounds[:-1], bounds[1:]):
        a2 = int(a + rng.integers(0, gap_max + 1))
        b2 = int(b - rng.integers(0, gap_max + 1))
        if b2 - a2 >= min_len:
            out.append((a2, b2))
    return out


@dataclass
class Segment:
    """One experiment: a truth segment, its simulated reads and a candidate contig set."""
    truth: bytes
    reads: np.ndarray            # uint8 [N, read_len] ASCII
    contigs: list                # list[bytes]
    contig_truth_start: list = field(default_factory=list)  # expected startpos (-1 if mutated)

    @property
    def read_list(self):
        return [r.tobytes() for r in self.reads]


def make_segment(seed, length=50_000, read_len=100, coverage=30.0, n_contigs=16,
                 prob8=None, min_contig=61, mut_frac=0.1, n_gap_scaffolds=0) -> Segment:
    rng = np.random.default_rng(seed)
    if prob8 is None:
        prob8 = tables.sub_table(tables.normalised(tables.load_raw()), 8)
    codes = random_truth_codes(rng, length)
    truth = codes_to_ascii(codes)
    starts = sample_read_starts(rng, codes, read_len, coverage, prob8)
    reads = truth[starts[:, None] + np.arange(read_len)[None, :]]
    contigs, where = [], []
    ivs = velvet_style_intervals(rng, length, n_contigs, min(min_contig, max(length // 4, 1)))
    for (a, b) in ivs:
        c = truth[a:b].copy()
        if rng.random() < mut_frac and b - a > 2:
            p = int(rng.integers(1, b - a - 1))
            c[p] = _ACGT[(np.searchsorted(_ACGT, c[p]) + 1 + int(rng.integers(0, 3))) % 4]
            where.append(-1)
        else:
            where.append(a)
        contigs.append(c.tobytes())
    for _ in range(n_gap_scaffolds):
        if len(ivs) >= 2:
            i, j = rng.choice(len(ivs), size=2, replace=False)
            gap = b"N" * int(rng.integers(5, 60))
            contigs.append(contigs[i] + gap + contigs[j])
            where.append(-1)
    return Segment(truth.tobytes(), reads, contigs, where)


@dataclass
class Batch:
    """Flat-buffer form of many segments (the layout the C-ABI takes, include/breakscore.h).

    String i of a set is ``chars[off[i]:off[i+1]]``; segment s owns reads
    ``seg_read_start[s]:seg_read_start[s+1]`` and contigs ``seg_contig_start[s]:...``.
    ``read_len`` > 0 means every read has that length and ``read_chars`` is dense
    (``read_off`` may then be omitted)."""
    read_chars: np.ndarray
    read_off: np.ndarray | None
    read_len: int
    contig_chars: np.ndarray
    contig_off: np.ndarray
    truth_chars: np.ndarray
    truth_off: np.ndarray
    seg_read_start: np.ndarray
    seg_contig_start: np.ndarray

    @property
    def n_segments(self):
        return len(self.truth_off) - 1

    @property
    def n_reads(self):
        return int(self.seg_read_start[-1])

    @property
    def n_contigs(self):
        return int(self.seg_contig_start[-1])

    def segment(self, s) -> Segment:
        r0, r1 = int(self.seg_read_start[s]), int(self.seg_read_start[s + 1])
        c0, c1 = int(self.seg_contig_start[s]), int(self.seg_contig_start[s + 1])
        if self.read_off is None:
            reads = self.read_chars[r0 * self.read_len:r1 * self.read_len].reshape(-1, self.read_len)
        else:
            raise ValueError("segment() view needs uniform read length")
        contigs = [self.contig_chars[self.contig_off[c]:self.contig_off[c + 1]].tobytes()
                   for c in range(c0, c1)]
        truth = self.truth_chars[self.truth_off[s]:self.truth_off[s + 1]].tobytes()
        return Segment(truth, reads, contigs)

    def pair_bases(self) -> float:
        """sum over segments of N_s * sum_c L_c (the 'read x contig bp compared' unit)."""
        n_s = np.diff(self.seg_read_start).astype(np.float64)
        lens = np.diff(self.contig_off).astype(np.float64)
        csum = np.concatenate([[0.0], np.cumsum(lens)])
        l_s = csum[self.seg_contig_start[1:]] - csum[self.seg_contig_start[:-1]]
        return float((n_s * l_s).sum())


def make_batch(n_segments, seed=1234, length=50_000, read_len=150, coverage=30.0,
               contigs_lo=5, contigs_hi=60, prob8=None, mut_frac=0.1, n_gap_scaffolds=0) -> Batch:
    """cfg-2 style study: n_segments independent segments, seeds seed+i."""
    if prob8 is None:
        prob8 = tables.sub_table(tables.normalised(tables.load_raw()), 8)
    rc, cc, tc = [], [], []
    c_lens, t_lens = [], []
    srs, scs = [0], [0]
    for i in range(n_segments):
        rng = np.random.default_rng(seed + i)
        nct = int(rng.integers(contigs_lo, contigs_hi + 1))
        seg = make_segment(seed + i, length, read_len, coverage, nct, prob8,
                           mut_frac=mut_frac, n_gap_scaffolds=n_gap_scaffolds)
        rc.append(seg.reads.reshape(-1))
        for c in seg.contigs:
            cc.append(np.frombuffer(c, dtype=np.uint8))
            c_lens.append(len(c))
        tc.append(np.frombuffer(seg.truth, dtype=np.uint8))
        t_lens.append(len(seg.truth))
        srs.append(srs[-1] + seg.reads.shape[0])
        scs.append(scs[-1] + len(seg.contigs))
    cat = lambda xs: np.concatenate(xs) if xs else np.zeros(0, np.uint8)
    off = lambda ls: np.concatenate([[0], np.cumsum(ls)]).astype(np.int64)
    return Batch(cat(rc), None, read_len, cat(cc), off(c_lens), cat(tc), off(t_lens),
                 np.asarray(srs, np.int64), np.asarray(scs, np.int64))


def make_scaffold_set(seed, length=50_000, read_len=150, coverage=30.0, n_base=16,
                      n_scaffolds=10_000, lo=10_000, hi=50_000, prob8=None) -> Segment:
    """cfg-4: candidates are concatenations of permutations of ~n_base base contigs, the
    shape upstream assemble_contigs (lib/BreakageScorer.cpp:105-171) produces."""
    rng = np.random.default_rng(seed)
    seg = make_segment(seed, length, read_len, coverage, n_base, prob8, mut_frac=0.0)
    base = [np.frombuffer(c, dtype=np.uint8) for c in seg.contigs]
    scaffolds, part_start, part_base = [], [0], []
    for _ in range(n_scaffolds):
        target = int(rng.integers(lo, hi + 1))
        order = rng.permutation(len(base))
        parts, tot = [], 0
        for j in order:
            parts.append(base[j])
            part_base.append(int(j))
            tot += len(base[j])
            if tot >= target:
                break
        part_start.append(len(part_base))
        scaffolds.append(np.concatenate(parts).tobytes())
    out = Segment(seg.truth, seg.reads, scaffolds)
    # how the candidates were put together (plain concatenations: every overlap 0), for bs_score_scaffolds
    out.base_contigs = list(seg.contigs)
    out.part_start = np.asarray(part_start, np.int64)
    out.part_base = np.asarray(part_base, np.int32)
    return out
