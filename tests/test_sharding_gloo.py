"""Multi-rank host logic on CPU: world_size-2 gloo processes shard a batch by segments, each
scores its share (with the oracle standing in for the GPU call -- this test is about the
partitioning and the gather, not about kernels) and rank 0's gathered records equal one
unsharded run."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from genomeassembler_dev_b200 import sharding, synth, tables


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _score_with_oracle(args_tuple, kmers, prob):
    from oracle import loader as O
    (rc, _, rl, cc, coff, tc, toff, srs, scs) = args_tuple
    res = {k: [] for k in sharding.RECORD_F64 + sharding.RECORD_I32}
    for s in range(len(toff) - 1):
        reads = [rc[n * rl:(n + 1) * rl].tobytes() for n in range(int(srs[s]), int(srs[s + 1]))]
        ctgs = [cc[coff[c]:coff[c + 1]].tobytes() for c in range(int(scs[s]), int(scs[s + 1]))]
        o = O.oracle_calc_breakscore(ctgs, reads, tc[toff[s]:toff[s + 1]].tobytes(), 8, kmers, prob, want_prob_dist=False)
        for k in res:
            res[k].append(o[k])
    return {k: np.concatenate(v) if v else np.zeros(0) for k, v in res.items()}


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    kmers, prob = tables.all_kmer_strings(), tables.normalised(tables.load_raw())
    b = synth.make_batch(5, seed=7, length=1500, read_len=30, coverage=6, contigs_lo=1, contigs_hi=4)
    runs = sharding.shard_segments(b.seg_read_start, b.seg_contig_start, b.contig_off, world)
    s0, s1 = runs[rank]
    counts = [int(b.seg_contig_start[e] - b.seg_contig_start[a]) for a, e in runs]
    res = _score_with_oracle(sharding.slice_batch(b, s0, s1), kmers, prob)
    rec = sharding.pack_records(res, counts[rank])
    got = sharding.gather_records(rec, counts)
    if rank == 0:
        whole = _score_with_oracle(sharding.slice_batch(b, 0, b.n_segments), kmers, prob)
        want = sharding.pack_records(whole, b.n_contigs)
        q.put(bool(np.array_equal(got, want, equal_nan=True)) and sum(counts) == b.n_contigs)
    dist.destroy_process_group()


def test_segment_sharding_and_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_shard_segments_covers_everything():
    b = synth.make_batch(9, seed=3, length=1200, read_len=25, coverage=5, contigs_lo=1, contigs_hi=3)
    for world in (1, 2, 4, 8, 16):
        runs = sharding.shard_segments(b.seg_read_start, b.seg_contig_start, b.contig_off, world)
        assert len(runs) == world and runs[0][0] == 0 and runs[-1][1] == b.n_segments
        assert all(runs[i][1] == runs[i + 1][0] for i in range(world - 1))


def test_lpt_contig_sharding():
    lens = np.array([50000, 100, 30000, 30000, 20, 999, 45000, 1])
    parts = sharding.shard_contigs_lpt(lens, 3)
    assert sorted(np.concatenate(parts).tolist()) == list(range(len(lens)))
    loads = [lens[p].sum() for p in parts]
    assert max(loads) - min(loads) <= lens.max()
    for p in parts:
        assert np.all(np.diff(p) > 0)
