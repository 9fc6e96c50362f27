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


def _contig_worker(rank, world, port, q, emul_lib):
    """cfg-4 shape: one segment, many candidate contigs, reads replicated, contigs LPT-sharded.  Each rank scores
    with the kernel sources under the CPU emulation (the product library needs a GPU)."""
    from genomeassembler_dev_b200 import breakscore
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    kmers, prob = tables.all_kmer_strings(), tables.normalised(tables.load_raw())
    b = synth.make_batch(1, seed=11, length=3000, read_len=40, coverage=8, contigs_lo=9, contigs_hi=9)
    ctgs = [b.contig_chars[b.contig_off[c]:b.contig_off[c + 1]].tobytes() for c in range(b.n_contigs)]
    reads = b.read_chars.reshape(-1, b.read_len)
    truth = b.truth_chars.tobytes()
    sc = breakscore.BreakageScorer(0, emul_lib)
    sc.set_table(kmers, prob)
    flags = breakscore.DEFAULT_FLAGS
    got, info = sharding.score_contigs_sharded(sc, ctgs, reads, truth, kmer=8, flags=flags)
    ok = len(info["local_index"]) == len(info["local"]["kmer_breaks"])
    if rank == 0:
        whole = sc.score(ctgs, reads, truth, kmer=8, flags=flags)
        for k in sharding.RECORD_F64 + sharding.RECORD_I32:
            ok = ok and bool(np.array_equal(np.asarray(got[k], dtype=np.float64), np.asarray(whole[k], dtype=np.float64), equal_nan=True))
        ok = ok and sorted(set(got["owner"].tolist())) == list(range(min(world, len(ctgs))))
        q.put(ok)
    else:
        ok = ok and got is None
        assert ok
    sc.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_contig_sharding_reads_replicated(world, emul_lib):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_contig_worker, args=(r, world, port, q, emul_lib)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def _read_shard_worker(rank, world, port, q, emul_lib):
    """cfg-5 shape: the READS sharded, position weights all-reduced (gloo here, NCCL on the GPUs), contig ranges scored
    from the sum.  Kernels under the CPU emulation."""
    from genomeassembler_dev_b200 import breakscore
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    kmers, prob = tables.all_kmer_strings(), tables.normalised(tables.load_raw())
    b = synth.make_batch(1, seed=12, length=3000, read_len=40, coverage=8, contigs_lo=7, contigs_hi=7)
    ctgs = [b.contig_chars[b.contig_off[c]:b.contig_off[c + 1]].tobytes() for c in range(b.n_contigs)]
    reads = b.read_chars.reshape(-1, b.read_len)
    truth = b.truth_chars.tobytes()
    sc = breakscore.BreakageScorer(0, emul_lib)
    sc.set_table(kmers, prob)
    n = len(reads)
    mine = reads[n * rank // world:n * (rank + 1) // world]
    got = sharding.score_reads_sharded(sc, ctgs, mine, truth, kmer=8, flags=breakscore.DEFAULT_FLAGS)
    if rank == 0:
        whole = sc.score(ctgs, reads, truth, kmer=8, flags=breakscore.DEFAULT_FLAGS)
        ok = all(np.array_equal(np.asarray(got[k], dtype=np.float64), np.asarray(whole[k], dtype=np.float64), equal_nan=True)
                 for k in sharding.RECORD_F64 + sharding.RECORD_I32)
        q.put(bool(ok))
    else:
        assert got is None
    sc.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_read_sharding_weights_all_reduced(world, emul_lib):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_read_shard_worker, args=(r, world, port, q, emul_lib)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def _scaffold_worker(rank, world, port, q, emul_lib):
    """a scaffold set given as parts, sharded like contigs: every rank scores its scaffolds from the parts (kernels under
    the CPU emulation); the gathered table equals one rank scoring the whole set, bit for bit"""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import scaffold_cases as SC
    from genomeassembler_dev_b200 import breakscore
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    kmers, prob = tables.all_kmer_strings(), tables.normalised(tables.load_raw())
    truth, reads, sset = SC.make_set(71, length=2500, read_len=40, coverage=8, n_base=7, n_scaffolds=11, overlap=9, lib_path=emul_lib)
    sc = breakscore.BreakageScorer(0, emul_lib)
    sc.set_table(kmers, prob)
    flags = breakscore.WANT_KS | breakscore.WANT_STARTPOS
    got, info = sharding.score_scaffolds_sharded(sc, sset, reads, truth, kmer=8, flags=flags)
    ok = len(info["local_index"]) == len(info["local"]["kmer_breaks"])
    if rank == 0:
        whole = sc.score_scaffolds(sset, reads, truth, kmer=8, flags=flags)
        texts = sc.score(sset.texts(), reads, truth, kmer=8, flags=flags)
        for k in sharding.RECORD_F64 + sharding.RECORD_I32:
            ok = ok and bool(np.array_equal(np.asarray(got[k], dtype=np.float64), np.asarray(whole[k], dtype=np.float64), equal_nan=True))
        for k in sharding.RECORD_I32 + ("ks_stat_prob_dist",):
            ok = ok and bool(np.array_equal(np.asarray(got[k], dtype=np.float64), np.asarray(texts[k], dtype=np.float64), equal_nan=True))
        ok = ok and sorted(set(got["owner"].tolist())) == list(range(min(world, len(sset))))
        q.put(ok)
    else:
        ok = ok and got is None
        assert ok
    sc.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_scaffold_set_sharding(world, emul_lib):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_scaffold_worker, args=(r, world, port, q, emul_lib)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
