"""Host-side mirror of the upstream scorer interface over the C-ABI (include/breakscore.h).

Upstream exports ``calc_breakscore(path, sequencing_reads, true_solution, kmer, bp_kmer,
bp_prob)`` from ``lib/BreakageScorer.cpp:185-191`` to R through Rcpp and returns a named list
(``:343-353``).  :func:`calc_breakscore` below has the same argument names, order and meaning
and returns a dict with the same member names (plus the KS statistics the R driver computes
from that list, ``lib/DeNovoAssembler.R:416-424``).  Everything is computed by
``libbreakscore.so`` (CUDA, sm_100a); this module only flattens strings into the flat-buffer
layout of the C-ABI.  There is no CPU path: if the library is missing or no B200 is visible
the call raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
# (BREAKSCORE_LIB: another build of the same library, e.g. a tuning variant made with `make OUT=... EXTRA_NVFLAGS=-D...`)
DEFAULT_LIB = os.environ.get("BREAKSCORE_LIB") or os.path.join(_PKG, "libbreakscore.so")

# flags (include/breakscore.h)
WANT_PROB_DIST = 0x001
WANT_KS = 0x002
WANT_HIST = 0x004
WANT_POS = 0x008
WANT_STARTPOS = 0x010
WANT_LEV = 0x020
WANT_SECOND_TABLE = 0x040
PLACE_SCAN = 0x100
PLACE_TILE = 0x800
DEVICE_CHARS = 0x200
DEVICE_RESULT = 0x400
WEIGHTS_OUT = 0x1000   # two-phase scoring (reads sharded over GPUs): stop after the placement, weights out
WEIGHTS_IN = 0x2000    # ... score from (summed) weights, no reads
DEFAULT_FLAGS = WANT_PROB_DIST | WANT_KS | WANT_STARTPOS

STAGES = ("h2d", "pack", "place", "score", "truth_spectrum", "prob_dist_ks", "ks_path_freq",
          "startpos", "d2h", "lev")

ABI_SYMBOLS = (
    "bs_abi_version", "bs_ctx_create", "bs_ctx_destroy", "bs_last_error", "bs_ctx_set_stream",
    "bs_ctx_synchronize", "bs_ctx_launch_count", "bs_ctx_enable_timing", "bs_ctx_last_timings",
    "bs_ctx_last_place_ms", "bs_ctx_set_poll", "bs_set_table", "bs_set_truth_table", "bs_set_second_table", "bs_score_batch", "bs_score", "bs_score_multi",
    "bs_host_alloc", "bs_host_free", "bs_assemble_contigs", "bs_assemble_last_error", "bs_string_list_size",
    "bs_string_list_bytes", "bs_string_list_copy", "bs_string_list_free", "bs_simulate_capacity", "bs_simulate_reads",
    "bs_score_scaffolds", "bs_scaffold_lengths", "bs_scaffold_texts", "bs_assemble_scaffolds", "bs_scaffold_list_size",
    "bs_scaffold_list_parts", "bs_scaffold_list_copy", "bs_scaffold_list_free",
)

POLL_FN = C.CFUNCTYPE(C.c_int, C.c_void_p)
ERR_INTERRUPTED = 7

_i64p = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)


class BreakscoreError(RuntimeError):
    """A non-zero status of the C-ABI (upstream analogue: Rcpp::stop)."""

    def __init__(self, code, text):
        super().__init__(f"breakscore error {code}: {text}")
        self.code = code


class _Batch(C.Structure):
    _fields_ = [
        ("n_segments", C.c_int64), ("n_reads", C.c_int64), ("n_contigs", C.c_int64),
        ("read_chars", C.c_void_p), ("read_off", C.c_void_p), ("read_len", C.c_int32),
        ("contig_chars", C.c_void_p), ("contig_off", C.c_void_p),
        ("truth_chars", C.c_void_p), ("truth_off", C.c_void_p),
        ("seg_read_start", C.c_void_p), ("seg_contig_start", C.c_void_p),
    ]


class _Result(C.Structure):
    _fields_ = [
        ("sequence_len", C.c_void_p), ("bp_score", C.c_void_p),
        ("bp_score_norm_by_break_freqs", C.c_void_p), ("bp_score_norm_by_len", C.c_void_p),
        ("kmer_breaks", C.c_void_p), ("path_prob_dist_startpos", C.c_void_p),
        ("lev_dist_vs_true", C.c_void_p), ("ks_stat_prob_dist", C.c_void_p),
        ("ks_stat_path_freq", C.c_void_p), ("path_prob_dist", C.c_void_p),
        ("path_prob_dist_off", C.c_void_p), ("hist", C.c_void_p), ("pos", C.c_void_p),
        ("pos_off", C.c_void_p),
        ("bp_score2", C.c_void_p), ("bp_score_norm_by_break_freqs2", C.c_void_p), ("bp_score_norm_by_len2", C.c_void_p),
        ("ks_stat_prob_dist2", C.c_void_p), ("ks_stat_path_freq2", C.c_void_p), ("path_prob_dist2", C.c_void_p),
        ("weights", C.c_void_p), ("weights_total", C.c_void_p),
    ]


class _ScaffoldSet(C.Structure):
    _fields_ = [
        ("n_base", C.c_int64), ("base_chars", C.c_void_p), ("base_off", C.c_void_p),
        ("n_scaffolds", C.c_int64), ("scaffold_part_start", C.c_void_p), ("part_base", C.c_void_p),
        ("part_overlap", C.c_void_p),
    ]


def load_library(path: str | None = None) -> C.CDLL:
    """dlopen the C-ABI library and declare the prototypes of include/breakscore.h."""
    path = path or DEFAULT_LIB
    if not os.path.exists(path):
        raise FileNotFoundError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(path)
    lib.bs_abi_version.restype = C.c_int
    lib.bs_ctx_create.restype = C.c_int
    lib.bs_ctx_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    lib.bs_ctx_destroy.restype = None
    lib.bs_ctx_destroy.argtypes = [C.c_void_p]
    lib.bs_last_error.restype = C.c_char_p
    lib.bs_last_error.argtypes = [C.c_void_p]
    lib.bs_ctx_set_stream.restype = C.c_int
    lib.bs_ctx_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    lib.bs_ctx_synchronize.restype = C.c_int
    lib.bs_ctx_synchronize.argtypes = [C.c_void_p]
    lib.bs_ctx_launch_count.restype = C.c_int64
    lib.bs_ctx_launch_count.argtypes = [C.c_void_p]
    lib.bs_ctx_enable_timing.restype = C.c_int
    lib.bs_ctx_enable_timing.argtypes = [C.c_void_p, C.c_int]
    lib.bs_ctx_last_timings.restype = C.c_int
    lib.bs_ctx_last_timings.argtypes = [C.c_void_p, _f64p, C.c_int]
    lib.bs_ctx_last_place_ms.restype = C.c_double
    lib.bs_ctx_last_place_ms.argtypes = [C.c_void_p]
    lib.bs_ctx_set_poll.restype = C.c_int
    lib.bs_ctx_set_poll.argtypes = [C.c_void_p, POLL_FN, C.c_void_p]
    lib.bs_set_table.restype = C.c_int
    lib.bs_set_table.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
    lib.bs_set_truth_table.restype = C.c_int
    lib.bs_set_truth_table.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    lib.bs_set_second_table.restype = C.c_int
    lib.bs_set_second_table.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    lib.bs_score_batch.restype = C.c_int
    lib.bs_score_batch.argtypes = [C.c_void_p, C.POINTER(_Batch), C.c_int, C.c_uint32, C.POINTER(_Result)]
    lib.bs_score.restype = C.c_int
    lib.bs_score.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                             C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_uint32, C.POINTER(_Result)]
    lib.bs_score_multi.restype = C.c_int
    lib.bs_score_multi.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                   C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_uint32, C.POINTER(_Result)]
    lib.bs_host_alloc.restype = C.c_void_p
    lib.bs_host_alloc.argtypes = [C.c_int64]
    lib.bs_host_free.restype = None
    lib.bs_host_free.argtypes = [C.c_void_p]
    lib.bs_assemble_contigs.restype = C.c_int
    lib.bs_assemble_contigs.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.POINTER(C.c_void_p)]
    lib.bs_assemble_last_error.restype = C.c_char_p
    lib.bs_string_list_size.restype = C.c_int64
    lib.bs_string_list_size.argtypes = [C.c_void_p]
    lib.bs_string_list_bytes.restype = C.c_int64
    lib.bs_string_list_bytes.argtypes = [C.c_void_p]
    lib.bs_string_list_copy.restype = None
    lib.bs_string_list_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.bs_string_list_free.restype = None
    lib.bs_string_list_free.argtypes = [C.c_void_p]
    lib.bs_simulate_capacity.restype = C.c_int64
    lib.bs_simulate_capacity.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_double]
    lib.bs_simulate_reads.restype = C.c_int
    lib.bs_simulate_reads.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_double, C.c_int,
                                      C.c_uint64, C.c_uint32, C.c_void_p, C.c_int64, C.c_void_p]
    lib.bs_score_scaffolds.restype = C.c_int
    lib.bs_score_scaffolds.argtypes = [C.c_void_p, C.POINTER(_ScaffoldSet), C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                       C.c_void_p, C.c_int64, C.c_int, C.c_uint32, C.POINTER(_Result)]
    lib.bs_scaffold_lengths.restype = C.c_int
    lib.bs_scaffold_lengths.argtypes = [C.POINTER(_ScaffoldSet), C.c_void_p]
    lib.bs_scaffold_texts.restype = C.c_int
    lib.bs_scaffold_texts.argtypes = [C.POINTER(_ScaffoldSet), C.c_void_p, C.c_void_p]
    lib.bs_assemble_scaffolds.restype = C.c_int
    lib.bs_assemble_scaffolds.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    lib.bs_scaffold_list_size.restype = C.c_int64
    lib.bs_scaffold_list_size.argtypes = [C.c_void_p]
    lib.bs_scaffold_list_parts.restype = C.c_int64
    lib.bs_scaffold_list_parts.argtypes = [C.c_void_p]
    lib.bs_scaffold_list_copy.restype = None
    lib.bs_scaffold_list_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.bs_scaffold_list_free.restype = None
    lib.bs_scaffold_list_free.argtypes = [C.c_void_p]
    if lib.bs_abi_version() != 5:
        raise RuntimeError(f"{path}: ABI version {lib.bs_abi_version()} != 5")
    return lib


def flatten(strings):
    """list of str / bytes -> (uint8 chars, int64 offsets[n+1])  (string i = chars[off[i]:off[i+1]])"""
    bs = [s.encode("ascii") if isinstance(s, str) else bytes(s) for s in strings]
    off = np.zeros(len(bs) + 1, dtype=np.int64)
    if bs:
        np.cumsum(np.fromiter((len(b) for b in bs), dtype=np.int64, count=len(bs)), out=off[1:])
    chars = np.frombuffer(b"".join(bs), dtype=np.uint8) if off[-1] else np.zeros(1, np.uint8)
    return chars, off


def prob_dist_offsets(contig_off, kmer):
    """where each contig's rolling-window vector goes: max(L_c - kmer + 1, 0) values each"""
    lens = np.maximum(np.diff(np.asarray(contig_off, dtype=np.int64)) - kmer + 1, 0)
    off = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    return off


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class ScaffoldSet:
    """Candidate scaffolds as PARTS of base contigs (include/breakscore.h: bs_scaffold_set): scaffold s is
    ``base[part_base[i]][part_overlap[i]:]`` joined over ``i in part_start[s]:part_start[s+1]``.  What upstream's
    ``assemble_contigs`` (lib/BreakageScorer.cpp:105-171) builds, before it is flattened into strings."""

    def __init__(self, base_contigs, part_start, part_base, part_overlap, lib_path: str | None = None):
        self.base_contigs = [c.encode("ascii") if isinstance(c, str) else bytes(c) for c in base_contigs]
        self.base_chars, self.base_off = flatten(self.base_contigs)
        self.part_start = np.ascontiguousarray(part_start, dtype=np.int64)
        self.part_base = np.ascontiguousarray(part_base, dtype=np.int32)
        self.part_overlap = np.ascontiguousarray(part_overlap, dtype=np.int32)
        self._lib_path = lib_path

    def __len__(self):
        return len(self.part_start) - 1

    def subset(self, index) -> "ScaffoldSet":
        """the scaffolds `index` (ascending input indices) as a set of their own over the same base contigs: what one
        rank scores when a set is sharded over GPUs (sharding.score_scaffolds_sharded)"""
        index = np.asarray(index, dtype=np.int64)
        counts = (self.part_start[1:] - self.part_start[:-1])[index]
        start = np.zeros(len(index) + 1, np.int64)
        np.cumsum(counts, out=start[1:])
        pick = np.concatenate([np.arange(self.part_start[i], self.part_start[i + 1]) for i in index]) if len(index) else np.zeros(0, np.int64)
        return ScaffoldSet(self.base_contigs, start, self.part_base[pick], self.part_overlap[pick], self._lib_path)

    def c_struct(self) -> _ScaffoldSet:
        return _ScaffoldSet(len(self.base_contigs), self.base_chars.ctypes.data, self.base_off.ctypes.data, len(self),
                            self.part_start.ctypes.data, self.part_base.ctypes.data, self.part_overlap.ctypes.data)

    def lengths(self) -> np.ndarray:
        lib = load_library(self._lib_path)
        out = np.zeros(max(len(self), 1), np.int64)
        rc = lib.bs_scaffold_lengths(C.byref(self.c_struct()), _ptr(out))
        if rc != 0:
            raise BreakscoreError(rc, lib.bs_assemble_last_error().decode())
        return out[:len(self)]

    def texts(self) -> list:
        """the scaffold strings (host; what upstream's assemble_contigs returns)"""
        lib = load_library(self._lib_path)
        off = np.zeros(len(self) + 1, np.int64)
        st = self.c_struct()
        rc = lib.bs_scaffold_texts(C.byref(st), None, _ptr(off))
        if rc != 0:
            raise BreakscoreError(rc, lib.bs_assemble_last_error().decode())
        chars = np.zeros(max(int(off[-1]), 1), np.uint8)
        rc = lib.bs_scaffold_texts(C.byref(st), _ptr(chars), _ptr(off))
        if rc != 0:
            raise BreakscoreError(rc, lib.bs_assemble_last_error().decode())
        return [chars[off[i]:off[i + 1]].tobytes() for i in range(len(self))]


class BreakageScorer:
    """A device context with a resident probability table (bs_ctx of include/breakscore.h)."""

    def __init__(self, device: int = 0, lib_path: str | None = None):
        self._lib = load_library(lib_path)
        self._ctx = C.c_void_p()
        rc = self._lib.bs_ctx_create(device, C.byref(self._ctx))
        if rc != 0:
            text = self._lib.bs_last_error(None).decode()
            self._ctx = C.c_void_p()
            raise BreakscoreError(rc, text)
        self.device = device
        self.n_table = 0
        self._table_key = None

    # -- lifetime ------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            self._lib.bs_ctx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc):
        if rc != 0:
            raise BreakscoreError(rc, self._lib.bs_last_error(self._ctx).decode())

    # -- context controls ----------------------------------------------------------------
    def set_stream(self, cuda_stream: int | None):
        self._check(self._lib.bs_ctx_set_stream(self._ctx, C.c_void_p(cuda_stream or 0)))

    def synchronize(self):
        self._check(self._lib.bs_ctx_synchronize(self._ctx))

    @property
    def launch_count(self) -> int:
        return int(self._lib.bs_ctx_launch_count(self._ctx))

    def set_poll(self, poll=None):
        """Interrupt poll for long calls: ``poll()`` is called on the calling thread between the pipeline chunks of a
        scoring call; a true return stops the call with BreakscoreError(code ERR_INTERRUPTED) -- the hook the Rcpp glue
        uses for ``Rcpp::checkUserInterrupt``.  ``None`` removes it."""
        self._poll_keepalive = POLL_FN((lambda _user: 1 if poll() else 0)) if poll is not None else C.cast(None, POLL_FN)
        self._check(self._lib.bs_ctx_set_poll(self._ctx, self._poll_keepalive, None))

    def enable_timing(self, on: bool = True):
        self._check(self._lib.bs_ctx_enable_timing(self._ctx, 1 if on else 0))

    def last_timings(self) -> dict:
        ms = np.zeros(len(STAGES), np.float64)
        n = self._lib.bs_ctx_last_timings(self._ctx, ms.ctypes.data_as(_f64p), len(STAGES))
        return {STAGES[i]: float(ms[i]) for i in range(n)}

    def pinned_empty(self, shape, dtype=np.uint8):
        """A numpy array in page-locked host memory (bs_host_alloc).  Inputs built in such arrays go to the device at the
        PCIe rate (~55 GB/s); pageable numpy memory is staged by the driver at about a fifth of that.  The memory is
        released when the array (and every view of it) is garbage collected."""
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        ptr = self._lib.bs_host_alloc(max(n, 1))
        if not ptr:
            raise MemoryError(f"bs_host_alloc({n}) failed")
        lib = self._lib

        class _Owner:  # frees the block when the last view goes away
            def __del__(self, p=ptr):
                lib.bs_host_free(p)

        buf = (C.c_ubyte * max(n, 1)).from_address(ptr)
        buf._owner = _Owner()
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    # -- table ---------------------------------------------------------------------------
    def set_table(self, bp_kmer, bp_prob, truth_prob=None):
        """bp_kmer / bp_prob of the upstream call (lib/BreakageScorer.cpp:194-197); truth_prob is
        the table behind the truth-side distribution of the KS statistics (the R driver keeps the
        real probabilities there in its "random" pass, lib/DeNovoAssembler.R:326-333)."""
        prob = np.ascontiguousarray(bp_prob, dtype=np.float64)
        if isinstance(bp_kmer, tuple) and len(bp_kmer) == 2 and isinstance(bp_kmer[0], np.ndarray):
            chars, off = bp_kmer
        else:
            chars, off = flatten(bp_kmer)
        if len(off) - 1 != len(prob):
            raise ValueError(f"bp_kmer has {len(off) - 1} rows, bp_prob {len(prob)}")
        self._check(self._lib.bs_set_table(self._ctx, _ptr(chars), _ptr(off), _ptr(prob), len(prob)))
        self.n_table = len(prob)
        if truth_prob is not None:
            t = np.ascontiguousarray(truth_prob, dtype=np.float64)
            self._check(self._lib.bs_set_truth_table(self._ctx, _ptr(t), len(t)))
        else:
            self._check(self._lib.bs_set_truth_table(self._ctx, None, 0))

    def set_second_table(self, prob2):
        """A second scoring table over the same rows (e.g. the uniform table of the R driver's "random"
        pass, lib/DeNovoAssembler.R:325-333); score with flags | WANT_SECOND_TABLE to get the ``*2`` members
        from the same placement.  None removes it."""
        if prob2 is None:
            self._check(self._lib.bs_set_second_table(self._ctx, None, 0))
            return
        t = np.ascontiguousarray(prob2, dtype=np.float64)
        self._check(self._lib.bs_set_second_table(self._ctx, _ptr(t), len(t)))

    # -- read simulation ------------------------------------------------------------------
    def simulate_reads(self, truths, read_len, coverage, seed=1234, kmer=8):
        """Reads of `read_len` bases sampled from every truth with upstream's law
        (lib/GenerateReads.R:302-313: starts drawn with replacement proportional to the table probability
        of the kmer-window, ceil(coverage * L / read_len) draws, overrunning draws dropped).
        Returns (uint8 [N, read_len], int64 seg_read_start[S+1]).  Uses the scoring table of the context."""
        tr, tr_off = flatten(truths)
        S = len(tr_off) - 1
        cap = int(self._lib.bs_simulate_capacity(_ptr(tr_off), S, int(read_len), float(coverage)))
        reads = np.zeros(max(cap, 1), np.uint8)
        srs = np.zeros(S + 1, np.int64)
        self._check(self._lib.bs_simulate_reads(self._ctx, _ptr(tr), _ptr(tr_off), S, int(read_len), float(coverage), int(kmer),
                                                int(seed) & 0xFFFFFFFFFFFFFFFF, 0, _ptr(reads), cap, _ptr(srs)))
        n = int(srs[-1])
        return reads[:n * read_len].reshape(n, read_len), srs

    # -- scoring -------------------------------------------------------------------------
    def score_batch(self, read_chars, read_off, read_len, contig_chars, contig_off, truth_chars,
                    truth_off, seg_read_start, seg_contig_start, kmer=8, flags=DEFAULT_FLAGS, group=None, weights=None,
                    scaffolds: "ScaffoldSet | None" = None):
        """Many independent segments in one call (one upstream calc_breakscore call each).
        Host numpy buffers in, dict of numpy arrays out (flat path_prob_dist + offsets).
        ``group``: more scorers (one per GPU, same tables) -- ONE segment's contigs are then dealt out over
        ``[self] + group`` by ``bs_score_multi`` (reads replicated, results in input order)."""
        contig_off = np.ascontiguousarray(contig_off, dtype=np.int64)
        truth_off = np.ascontiguousarray(truth_off, dtype=np.int64)
        srs = np.ascontiguousarray(seg_read_start, dtype=np.int64)
        scs = np.ascontiguousarray(seg_contig_start, dtype=np.int64)
        read_chars = np.ascontiguousarray(read_chars, dtype=np.uint8)
        contig_chars = np.ascontiguousarray(contig_chars, dtype=np.uint8)
        truth_chars = np.ascontiguousarray(truth_chars, dtype=np.uint8)
        if read_off is not None:
            read_off = np.ascontiguousarray(read_off, dtype=np.int64)
        S, N, Cn = len(truth_off) - 1, int(srs[-1]), int(scs[-1])
        b = _Batch(S, N, Cn, read_chars.ctypes.data, None if read_off is None else read_off.ctypes.data,
                   int(read_len or 0), contig_chars.ctypes.data, contig_off.ctypes.data,
                   truth_chars.ctypes.data, truth_off.ctypes.data, srs.ctypes.data, scs.ctypes.data)
        out = {
            "sequence_len": np.zeros(Cn, np.int32),
            "bp_score": np.zeros(Cn, np.float64),
            "bp_score_norm_by_break_freqs": np.zeros(Cn, np.float64),
            "bp_score_norm_by_len": np.zeros(Cn, np.float64),
            "kmer_breaks": np.zeros(Cn, np.int32),
            "lev_dist_vs_true": np.zeros(Cn, np.int32),
            "path_prob_dist_startpos": np.zeros(Cn, np.int32),
        }
        r = _Result()
        for k, v in out.items():
            setattr(r, k, v.ctypes.data)
        keep = [read_chars, contig_chars, truth_chars, read_off, contig_off, truth_off, srs, scs]
        if flags & WANT_KS:
            out["ks_stat_prob_dist"] = np.zeros(Cn, np.float64)
            out["ks_stat_path_freq"] = np.zeros(Cn, np.float64)
            r.ks_stat_prob_dist = out["ks_stat_prob_dist"].ctypes.data
            r.ks_stat_path_freq = out["ks_stat_path_freq"].ctypes.data
        if flags & WANT_PROB_DIST:
            pd_off = prob_dist_offsets(contig_off, kmer)
            out["path_prob_dist_flat"] = np.zeros(max(int(pd_off[-1]), 1), np.float64)
            out["path_prob_dist_off"] = pd_off
            r.path_prob_dist = out["path_prob_dist_flat"].ctypes.data
            r.path_prob_dist_off = pd_off.ctypes.data
        if flags & WANT_SECOND_TABLE:
            for k in ("bp_score2", "bp_score_norm_by_break_freqs2", "bp_score_norm_by_len2"):
                out[k] = np.zeros(Cn, np.float64)
                setattr(r, k, out[k].ctypes.data)
            if flags & WANT_KS:
                for k in ("ks_stat_prob_dist2", "ks_stat_path_freq2"):
                    out[k] = np.zeros(Cn, np.float64)
                    setattr(r, k, out[k].ctypes.data)
            if flags & WANT_PROB_DIST:
                out["path_prob_dist2_flat"] = np.zeros(max(int(pd_off[-1]), 1), np.float64)
                r.path_prob_dist2 = out["path_prob_dist2_flat"].ctypes.data
        if flags & WANT_HIST:
            out["hist"] = np.zeros((Cn, self.n_table + 1), np.int32)
            r.hist = out["hist"].ctypes.data
        if flags & WANT_POS:
            nr_of_contig = np.repeat(np.diff(srs), np.diff(scs))
            pos_off = np.zeros(Cn + 1, np.int64)
            np.cumsum(nr_of_contig, out=pos_off[1:])
            out["pos_flat"] = np.zeros(max(int(pos_off[-1]), 1), np.int32)
            out["pos_off"] = pos_off
            r.pos = out["pos_flat"].ctypes.data
            r.pos_off = pos_off.ctypes.data
        if weights is not None:
            r.weights, r.weights_total = weights
        if scaffolds is not None:
            # the contigs are a scaffold set given as parts: contig_chars is not used, contig_off = its text offsets
            if S != 1 or group:
                raise ValueError("a scaffold set is scored for ONE segment on one scorer")
            st = scaffolds.c_struct()
            self._check(self._lib.bs_score_scaffolds(self._ctx, C.byref(st), read_chars.ctypes.data,
                                                     None if read_off is None else read_off.ctypes.data, N, int(read_len or 0),
                                                     truth_chars.ctypes.data + int(truth_off[0]), int(truth_off[1] - truth_off[0]),
                                                     int(kmer), int(flags), C.byref(r)))
        elif group:
            if S != 1:
                raise ValueError("a scorer group shards the contigs of ONE segment (shard whole segments over processes instead)")
            if read_off is None:
                read_off = np.arange(N + 1, dtype=np.int64) * int(read_len or 0)
            ctxs = (C.c_void_p * (1 + len(group)))(self._ctx, *[g._ctx for g in group])
            self._check(self._lib.bs_score_multi(ctxs, len(ctxs), contig_chars.ctypes.data, contig_off.ctypes.data, Cn,
                                                 read_chars.ctypes.data, read_off.ctypes.data, N, truth_chars.ctypes.data + int(truth_off[0]),
                                                 int(truth_off[1] - truth_off[0]), int(kmer), int(flags), C.byref(r)))
        else:
            self._check(self._lib.bs_score_batch(self._ctx, C.byref(b), int(kmer), int(flags), C.byref(r)))
        del keep
        return out

    # -- two-phase scoring: the READS of one segment sharded over several GPUs (sharding.score_reads_sharded) --------
    def place_weights(self, path, sequencing_reads, weights_ptr: int, weights_total_ptr: int, flags=0):
        """Phase 1: place THIS rank's reads (uint8 [n, L] or a list of strings) in ALL contigs and leave the position
        weights in the caller's DEVICE arrays (int32: sum(L_c) + C entries / C entries; see include/breakscore.h).
        The weights of disjoint read sets add: sum them over the ranks, then call :meth:`score_from_weights`."""
        ct, ct_off = flatten(path)
        if isinstance(sequencing_reads, np.ndarray) and sequencing_reads.ndim == 2:
            rd = np.ascontiguousarray(sequencing_reads, dtype=np.uint8).reshape(-1)
            rd_off, rlen, nr = None, sequencing_reads.shape[1], sequencing_reads.shape[0]
        else:
            rd, rd_off = flatten(sequencing_reads)
            rlen, nr = 0, len(sequencing_reads)
        tr_off = np.zeros(2, np.int64)
        srs, scs = np.array([0, nr], np.int64), np.array([0, len(path)], np.int64)
        tr = np.zeros(1, np.uint8)
        b = _Batch(1, nr, len(path), rd.ctypes.data, None if rd_off is None else rd_off.ctypes.data, int(rlen), ct.ctypes.data,
                   ct_off.ctypes.data, tr.ctypes.data, tr_off.ctypes.data, srs.ctypes.data, scs.ctypes.data)
        r = _Result()
        r.weights, r.weights_total = weights_ptr, weights_total_ptr
        self._check(self._lib.bs_score_batch(self._ctx, C.byref(b), 8, int(flags) | WEIGHTS_OUT, C.byref(r)))

    def score_from_weights(self, path, true_solution, weights_ptr: int, weights_total_ptr: int, kmer=8, flags=DEFAULT_FLAGS):
        """Phase 2: score the contigs `path` from (summed) position weights; same dict as :meth:`score`."""
        ct, ct_off = flatten(path)
        tr, tr_off = flatten([true_solution])
        empty = np.zeros(1, np.uint8)
        return self._score_prepared(empty, None, 1, ct, ct_off, tr, tr_off, [0, 0], [0, len(path)], kmer, int(flags) | WEIGHTS_IN, path,
                                    weights=(weights_ptr, weights_total_ptr))

    def _score_prepared(self, rd, rd_off, rlen, ct, ct_off, tr, tr_off, srs, scs, kmer, flags, path, weights=None):
        res = self.score_batch(rd, rd_off, rlen, ct, ct_off, tr, tr_off, srs, scs, kmer=kmer, flags=flags, weights=weights)
        res["sequence"] = list(path)
        if flags & WANT_PROB_DIST:
            flat, off = res.pop("path_prob_dist_flat"), res.pop("path_prob_dist_off")
            res["path_prob_dist"] = [flat[off[i]:off[i + 1]] for i in range(len(path))]
        return res

    def score_scaffolds(self, scaffolds: ScaffoldSet, sequencing_reads, true_solution, kmer=8, flags=DEFAULT_FLAGS):
        """:meth:`score` for a candidate set given as parts of base contigs (``assemble_scaffolds``): the same dict, but the
        reads are placed once per BASE contig and every scaffold is scored from its parts and junction windows
        (``bs_score_scaffolds``); the scaffold texts never leave the device.  ``res["sequence"]`` holds the texts."""
        lens = scaffolds.lengths()
        ct_off = np.zeros(len(scaffolds) + 1, np.int64)
        np.cumsum(lens, out=ct_off[1:])
        if isinstance(sequencing_reads, np.ndarray) and sequencing_reads.ndim == 2:
            rd = np.ascontiguousarray(sequencing_reads, dtype=np.uint8).reshape(-1)
            rd_off, rlen, nr = None, sequencing_reads.shape[1], sequencing_reads.shape[0]
        else:
            rd, rd_off = flatten(sequencing_reads)
            rlen, nr = 0, len(sequencing_reads)
        tr, tr_off = flatten([true_solution])
        n = len(scaffolds)
        res = self.score_batch(rd, rd_off, rlen, np.zeros(1, np.uint8), ct_off, tr, tr_off, [0, nr], [0, n], kmer=kmer, flags=flags,
                               scaffolds=scaffolds)
        res["sequence"] = scaffolds.texts()
        if flags & WANT_PROB_DIST:
            flat, off = res.pop("path_prob_dist_flat"), res.pop("path_prob_dist_off")
            res["path_prob_dist"] = [flat[off[i]:off[i + 1]] for i in range(n)]
            if flags & WANT_SECOND_TABLE:
                flat2 = res.pop("path_prob_dist2_flat")
                res["path_prob_dist2"] = [flat2[off[i]:off[i + 1]] for i in range(n)]
        if flags & WANT_POS:
            flat = res.pop("pos_flat")
            res.pop("pos_off")
            res["pos"] = flat[:n * nr].reshape(n, nr)
        return res

    def score_batch_raw(self, batch: _Batch, result: _Result, kmer: int, flags: int):
        """Thin call for callers that manage their own (possibly device) buffers: bench.py."""
        self._check(self._lib.bs_score_batch(self._ctx, C.byref(batch), int(kmer), int(flags), C.byref(result)))

    def score(self, path, sequencing_reads, true_solution, kmer=8, flags=DEFAULT_FLAGS, group=None):
        """One segment; returns the upstream list (lib/BreakageScorer.cpp:343-353) as a dict.
        ``group``: further scorers (one per GPU) to share the contigs with, see :meth:`score_batch`."""
        ct, ct_off = flatten(path)
        if isinstance(sequencing_reads, np.ndarray) and sequencing_reads.ndim == 2:
            rd = np.ascontiguousarray(sequencing_reads, dtype=np.uint8).reshape(-1)
            rd_off, rlen, nr = None, sequencing_reads.shape[1], sequencing_reads.shape[0]
        else:
            rd, rd_off = flatten(sequencing_reads)
            rlen, nr = 0, len(sequencing_reads)
        tr, tr_off = flatten([true_solution])
        res = self.score_batch(rd, rd_off, rlen, ct, ct_off, tr, tr_off, [0, nr], [0, len(path)],
                               kmer=kmer, flags=flags, group=group)
        res["sequence"] = list(path)
        if flags & WANT_PROB_DIST:
            flat, off = res.pop("path_prob_dist_flat"), res.pop("path_prob_dist_off")
            res["path_prob_dist"] = [flat[off[i]:off[i + 1]] for i in range(len(path))]
            if flags & WANT_SECOND_TABLE:
                flat2 = res.pop("path_prob_dist2_flat")
                res["path_prob_dist2"] = [flat2[off[i]:off[i + 1]] for i in range(len(path))]
        if flags & WANT_POS:
            flat = res.pop("pos_flat")
            res.pop("pos_off")
            res["pos"] = flat[:len(path) * nr].reshape(len(path), nr)
        return res


_default = {}


def default_scorer(device: int = 0, lib_path: str | None = None) -> BreakageScorer:
    """Process-lifetime context per device (the Rcpp glue keeps the same kind of singleton)."""
    key = (device, lib_path or DEFAULT_LIB)
    if key not in _default:
        _default[key] = BreakageScorer(device, lib_path)
    return _default[key]


def calc_breakscore(path, sequencing_reads, true_solution, kmer, bp_kmer, bp_prob, *,
                    truth_prob=None, flags=DEFAULT_FLAGS, device=0, lib_path=None):
    """Drop-in for the upstream export (lib/BreakageScorer.cpp:185-191), same argument list.

    Returns a dict with the upstream list members ``sequence, sequence_len, bp_score,
    bp_score_norm_by_break_freqs, bp_score_norm_by_len, kmer_breaks, lev_dist_vs_true,
    path_prob_dist_startpos, path_prob_dist`` and, with WANT_KS, ``ks_stat_prob_dist`` /
    ``ks_stat_path_freq`` (the statistic of lib/DeNovoAssembler.R:416-424 for the two variants
    of the per-contig vector).  ``lev_dist_vs_true`` is computed with WANT_LEV (infix edit distance, the
    semantics of upstream's edlib call), else 0.
    """
    sc = default_scorer(device, lib_path)
    prob = np.ascontiguousarray(bp_prob, dtype=np.float64)
    tprob = None if truth_prob is None else np.ascontiguousarray(truth_prob, dtype=np.float64)
    names = ",".join(k if isinstance(k, str) else bytes(k).decode("ascii") for k in bp_kmer)
    key = (names, prob.tobytes(), None if tprob is None else tprob.tobytes())
    if sc._table_key != key:
        sc.set_table(bp_kmer, prob, tprob)
        sc._table_key = key
    return sc.score(path, sequencing_reads, true_solution, kmer=kmer, flags=flags)


def assemble_contigs(velvet_contigs, dbg_kmer, seed, *, n_shuffles=20000, n_threads=0, lib_path=None):
    """Drop-in for the upstream export ``assemble_contigs(velvet_contigs, dbg_kmer, seed)``
    (lib/BreakageScorer.cpp:79-83): the scaffold explosion that produces the candidate set scored by
    :func:`calc_breakscore`.  Host code in the same library (all cores); returns a list of bytes,
    longest first, in upstream's order."""
    lib = load_library(lib_path)
    chars, off = flatten(velvet_contigs)
    h = C.c_void_p()
    rc = lib.bs_assemble_contigs(_ptr(chars), _ptr(off), len(velvet_contigs), int(dbg_kmer), int(seed), int(n_shuffles),
                                 int(n_threads), C.byref(h))
    if rc != 0:
        raise BreakscoreError(rc, lib.bs_assemble_last_error().decode())
    try:
        n = lib.bs_string_list_size(h)
        out_chars = np.zeros(max(int(lib.bs_string_list_bytes(h)), 1), np.uint8)
        out_off = np.zeros(n + 1, np.int64)
        lib.bs_string_list_copy(h, _ptr(out_chars), _ptr(out_off))
    finally:
        lib.bs_string_list_free(h)
    return [out_chars[out_off[i]:out_off[i + 1]].tobytes() for i in range(n)]


def assemble_scaffolds(velvet_contigs, dbg_kmer, seed, *, n_shuffles=20000, n_threads=0, lib_path=None):
    """:func:`assemble_contigs` that keeps HOW every scaffold was glued together: returns ``(strings, ScaffoldSet)`` -- the
    same strings in the same (upstream) order, and their parts for :meth:`BreakageScorer.score_scaffolds`."""
    lib = load_library(lib_path)
    chars, off = flatten(velvet_contigs)
    h, hp = C.c_void_p(), C.c_void_p()
    rc = lib.bs_assemble_scaffolds(_ptr(chars), _ptr(off), len(velvet_contigs), int(dbg_kmer), int(seed), int(n_shuffles),
                                   int(n_threads), C.byref(h), C.byref(hp))
    if rc != 0:
        raise BreakscoreError(rc, lib.bs_assemble_last_error().decode())
    try:
        n = lib.bs_string_list_size(h)
        out_chars = np.zeros(max(int(lib.bs_string_list_bytes(h)), 1), np.uint8)
        out_off = np.zeros(n + 1, np.int64)
        lib.bs_string_list_copy(h, _ptr(out_chars), _ptr(out_off))
        n_parts = int(lib.bs_scaffold_list_parts(hp))
        part_start = np.zeros(n + 1, np.int64)
        part_base = np.zeros(max(n_parts, 1), np.int32)
        part_overlap = np.zeros(max(n_parts, 1), np.int32)
        lib.bs_scaffold_list_copy(hp, _ptr(part_start), _ptr(part_base), _ptr(part_overlap))
    finally:
        lib.bs_string_list_free(h)
        lib.bs_scaffold_list_free(hp)
    strings = [out_chars[out_off[i]:out_off[i + 1]].tobytes() for i in range(n)]
    return strings, ScaffoldSet(velvet_contigs, part_start, part_base[:n_parts], part_overlap[:n_parts], lib_path)
