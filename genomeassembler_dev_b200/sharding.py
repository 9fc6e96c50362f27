"""Partitioning of a scoring job over the GPUs of one box (one process per GPU).

Segments and contigs are independent units of the path (upstream resets all per-contig state at
``lib/BreakageScorer.cpp:279-297``), so a rank scores its share with no data-path collective and the
fixed-width per-contig records are gathered afterwards.

* many segments (cfg-2 study): contiguous runs of whole segments balanced by read x contig work;
* one segment, many contigs (cfg-4 scaffold sets, cfg-5): contigs dealt out longest-first (LPT) with
  the segment's reads replicated to every rank; a scaffold set given as parts of base contigs the same way, every rank
  scoring its share from the parts.
"""
from __future__ import annotations

import numpy as np


def shard_segments(seg_read_start, seg_contig_start, contig_off, world: int):
    """Split segments [0, S) into `world` contiguous runs of about equal sum(N_s * L_s) work.
    Returns a list of (s0, s1) per rank (possibly empty runs when S < world)."""
    srs = np.asarray(seg_read_start, dtype=np.int64)
    scs = np.asarray(seg_contig_start, dtype=np.int64)
    coff = np.asarray(contig_off, dtype=np.int64)
    S = len(srs) - 1
    n_s = np.diff(srs).astype(np.float64)
    l_s = (coff[scs[1:]] - coff[scs[:-1]]).astype(np.float64)
    work = n_s * l_s + 1.0
    cum = np.concatenate([[0.0], np.cumsum(work)])
    bounds = [0]
    for r in range(1, world):
        bounds.append(int(np.searchsorted(cum, cum[-1] * r / world, side="left")))
    bounds.append(S)
    bounds = np.maximum.accumulate(np.minimum(bounds, S))
    return [(int(bounds[r]), int(bounds[r + 1])) for r in range(world)]


def shard_contigs_lpt(contig_lens, world: int):
    """Longest-processing-time-first assignment of contigs to ranks.  Returns a list of index arrays
    (ascending inside a rank, so a rank's results stay in input order)."""
    lens = np.asarray(contig_lens, dtype=np.int64)
    order = np.argsort(-lens, kind="stable")
    load = np.zeros(world, dtype=np.int64)
    owner = np.zeros(len(lens), dtype=np.int64)
    for i in order:
        r = int(np.argmin(load))
        owner[i] = r
        load[r] += int(lens[i]) + 1
    return [np.nonzero(owner == r)[0] for r in range(world)]


def slice_batch(batch, s0: int, s1: int):
    """The sub-batch of segments [s0, s1) of a synth.Batch-like object (uniform read length), as the
    argument tuple of BreakageScorer.score_batch."""
    r0, r1 = int(batch.seg_read_start[s0]), int(batch.seg_read_start[s1])
    c0, c1 = int(batch.seg_contig_start[s0]), int(batch.seg_contig_start[s1])
    rl = batch.read_len
    return (batch.read_chars[r0 * rl:r1 * rl], None, rl,
            batch.contig_chars[batch.contig_off[c0]:batch.contig_off[c1]], batch.contig_off[c0:c1 + 1] - batch.contig_off[c0],
            batch.truth_chars[batch.truth_off[s0]:batch.truth_off[s1]], batch.truth_off[s0:s1 + 1] - batch.truth_off[s0],
            batch.seg_read_start[s0:s1 + 1] - r0, batch.seg_contig_start[s0:s1 + 1] - c0)


RECORD_F64 = ("bp_score", "bp_score_norm_by_break_freqs", "bp_score_norm_by_len", "ks_stat_prob_dist", "ks_stat_path_freq")
RECORD_I32 = ("sequence_len", "kmer_breaks", "path_prob_dist_startpos")


def pack_records(res: dict, n: int) -> np.ndarray:
    """Fixed-width per-contig record block [n, 8] float64 (ints are exact in float64) for the gather."""
    out = np.full((n, len(RECORD_F64) + len(RECORD_I32)), np.nan, dtype=np.float64)
    for j, k in enumerate(RECORD_F64):
        if k in res:
            out[:, j] = res[k]
    for j, k in enumerate(RECORD_I32):
        out[:, len(RECORD_F64) + j] = res[k]
    return out


def unpack_records(rec: np.ndarray) -> dict:
    out = {k: rec[:, j].copy() for j, k in enumerate(RECORD_F64)}
    for j, k in enumerate(RECORD_I32):
        out[k] = rec[:, len(RECORD_F64) + j].astype(np.int32)
    return out


def gather_records(rec: np.ndarray, counts, group=None, dst: int = 0):
    """torch.distributed gather of per-rank record blocks (variable row counts) to rank `dst`.
    Works with NCCL (CUDA tensors) and gloo (CPU tensors, used by the CPU tests)."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    width = rec.shape[1]
    pad = int(max(counts))
    mine = torch.full((pad, width), float("nan"), dtype=torch.float64, device=dev)
    if rec.shape[0]:
        mine[:rec.shape[0]] = torch.from_numpy(rec).to(dev)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    if rank != dst:
        return None
    return np.concatenate([out[r][:int(counts[r])].cpu().numpy() for r in range(world)], axis=0)


def score_contigs_sharded(scorer, path, sequencing_reads, true_solution, kmer=8, flags=None, group=None, dst: int = 0):
    """One segment with many candidate contigs / scaffolds (cfg-4, cfg-5) over the ranks of a process group:
    contigs are dealt out longest-first (:func:`shard_contigs_lpt`), every rank places the WHOLE read set
    (replicated input, as BASELINE.json's north_star has it) against its own contigs with `scorer`
    (a :class:`breakscore.BreakageScorer` whose table is set), the fixed-width score records are gathered
    and rank `dst` gets them back in INPUT order -- the order of the upstream list
    (``lib/BreakageScorer.cpp:306-353``; its sort is disabled at ``:311-315``).  Per-contig results do not depend on
    which rank scored them (fixed reduction order), so the gathered table is bit-identical to an unsharded call.

    Returns on rank `dst` a dict of the record columns (RECORD_F64 + RECORD_I32) plus ``owner`` (rank per contig);
    ``None`` elsewhere.  Every rank also gets ``local`` -- its own full result dict (``path_prob_dist`` etc. stay on
    the rank that computed them: the variable-length vectors are not sent through the collective) -- and
    ``local_index`` via the second return value."""
    import torch.distributed as dist

    from . import breakscore

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lens = np.fromiter((len(p) for p in path), dtype=np.int64, count=len(path))
    parts = shard_contigs_lpt(lens, world)
    mine = parts[rank]
    counts = [len(p) for p in parts]
    if flags is None:
        flags = breakscore.DEFAULT_FLAGS
    if len(mine):
        local = scorer.score([path[i] for i in mine], sequencing_reads, true_solution, kmer=kmer, flags=flags)
    else:  # more ranks than contigs: nothing to score here, but the collective still needs this rank
        local = {k: np.zeros(0) for k in RECORD_F64 + RECORD_I32}
    rec = pack_records(local, len(mine))
    got = gather_records(rec, counts, group=group, dst=dst)
    local_info = {"local": local, "local_index": mine}
    if rank != dst:
        return None, local_info
    order = np.concatenate(parts) if len(path) else np.zeros(0, np.int64)
    back = np.empty((len(path), rec.shape[1]), dtype=np.float64)
    back[order] = got
    out = unpack_records(back)
    owner = np.empty(len(path), dtype=np.int32)
    for r, p in enumerate(parts):
        owner[p] = r
    out["owner"] = owner
    return out, local_info


def score_scaffolds_sharded(scorer, scaffolds, sequencing_reads, true_solution, kmer=8, flags=None, group=None, dst: int = 0):
    """:func:`score_contigs_sharded` for a candidate set given as PARTS of base contigs (``breakscore.ScaffoldSet``, what
    ``assemble_scaffolds`` returns): the scaffolds are dealt out longest-first, every rank gets all base contigs and the
    whole read set and scores its share from the parts (``bs_score_scaffolds``: reads placed once per base contig on
    every rank -- that part is replicated, it is the small one), records gathered into input order on rank `dst`.
    Bit-identical to one rank scoring the whole set (the per-scaffold reductions have a fixed order)."""
    import torch.distributed as dist

    from . import breakscore

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lens = scaffolds.lengths()
    parts = shard_contigs_lpt(lens, world)
    mine = parts[rank]
    counts = [len(p) for p in parts]
    if flags is None:
        flags = breakscore.WANT_KS | breakscore.WANT_STARTPOS
    if len(mine):
        local = scorer.score_scaffolds(scaffolds.subset(mine), sequencing_reads, true_solution, kmer=kmer, flags=flags)
    else:
        local = {k: np.zeros(0) for k in RECORD_F64 + RECORD_I32}
    rec = pack_records(local, len(mine))
    got = gather_records(rec, counts, group=group, dst=dst)
    local_info = {"local": local, "local_index": mine}
    if rank != dst:
        return None, local_info
    order = np.concatenate(parts) if len(lens) else np.zeros(0, np.int64)
    back = np.empty((len(lens), rec.shape[1]), dtype=np.float64)
    back[order] = got
    out = unpack_records(back)
    owner = np.empty(len(lens), dtype=np.int32)
    for r, p in enumerate(parts):
        owner[p] = r
    out["owner"] = owner
    return out, local_info


def balanced_ranges(contig_lens, world: int):
    """[0, C) cut into `world` contiguous ranges of about equal total length; returns world + 1 bounds"""
    lens = np.asarray(contig_lens, dtype=np.int64)
    cum = np.concatenate([[0], np.cumsum(lens + 1)])
    bounds = [0]
    for r in range(1, world):
        bounds.append(int(np.searchsorted(cum, cum[-1] * r / world, side="left")))
    bounds.append(len(lens))
    return [int(x) for x in np.maximum.accumulate(np.minimum(bounds, len(lens)))]


def score_reads_sharded(scorer, path, reads_shard, true_solution, kmer=8, flags=None, group=None, dst: int = 0):
    """One segment whose READS are sharded over the ranks (cfg-5: the read set is the big input, and packing +
    indexing it is most of the work, so replicating it does not scale).  Upstream places every read on its own
    (``lib/BreakageScorer.cpp:235-243``), so the per-position break counts of disjoint read sets ADD:

      phase 1  every rank places ITS reads in ALL contigs            -> position weights on its device
      exchange all-reduce of the weights over the group (NCCL on GPUs; the path's one real data exchange)
      phase 2  every rank scores a contiguous range of the contigs from the summed weights; records gathered

    Returns on rank `dst` the record columns in input order (bit-identical to one unsharded call), None elsewhere."""
    import torch
    import torch.distributed as dist

    from . import breakscore

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    if flags is None:
        flags = breakscore.DEFAULT_FLAGS
    lens = np.fromiter((len(p) for p in path), dtype=np.int64, count=len(path))
    off = np.concatenate([[0], np.cumsum(lens)])
    C = len(path)
    w = torch.zeros(int(off[-1]) + C + 1, dtype=torch.int32, device=dev)
    tot = torch.zeros(C + 1, dtype=torch.int32, device=dev)
    scorer.place_weights(path, reads_shard, w.data_ptr(), tot.data_ptr())
    scorer.synchronize()
    dist.all_reduce(w, group=group)
    dist.all_reduce(tot, group=group)
    if dev.type == "cuda":
        torch.cuda.synchronize()
    bounds = balanced_ranges(lens, world)
    c0, c1 = bounds[rank], bounds[rank + 1]
    counts = [bounds[r + 1] - bounds[r] for r in range(world)]
    if c1 > c0:
        w_ptr = w.data_ptr() + 4 * (int(off[c0]) + c0)
        local = scorer.score_from_weights(path[c0:c1], true_solution, w_ptr, tot.data_ptr() + 4 * c0, kmer=kmer, flags=flags)
    else:
        local = {k: np.zeros(0) for k in RECORD_F64 + RECORD_I32}
    got = gather_records(pack_records(local, c1 - c0), counts, group=group, dst=dst)
    if rank != dst:
        return None
    return unpack_records(got)
