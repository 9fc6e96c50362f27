"""BASELINE.json configs[3] and configs[4] with north_star's multi-GPU split, measured inside bench.py's run:
ONE segment, its contigs dealt out over the ranks longest-first (sharding.shard_contigs_lpt), the reads and the truth
replicated on every rank, the fixed-width score records all-gathered over NCCL and put back into input order
(upstream keeps input order: lib/BreakageScorer.cpp:306-353).

  cfg4       10 000 candidate scaffolds of 10-50 kb (concatenations of ~16 base contigs, the shape upstream's
             assemble_contigs produces: lib/BreakageScorer.cpp:105-171) against one 50 kb segment's reads
  cfg5       100 Mb truth, 1e8 reads of 150 bp, 1e5 contigs of ~1 kb (5 % with one substitution), generated on the
             device (uniform read starts: SURVEY.md appendix C allows it for this config)
  cfg5_mini  the same shape at 1 Mb / 3e5 reads / 1000 contigs: small enough for the CPU oracle to see ALL reads

Per config: device-resident ms per step (CUDA events, max over ranks), host-to-host ms per step two ways (every rank
copies all the reads in / every rank copies 1/N of them in and NCCL all-gathers the rest over NVLink), a sampled oracle
diff per rank (>= 32 contigs, all reads: integer columns bit-exact, fp64 within 1e-9), the gathered table compared with
ONE GPU scoring every contig alone (bit-identical: per-contig reductions have a fixed order), and -- cfg5 at full size,
where no CPU oracle can see 1e8 reads -- kmer_breaks and startpos of ALL contigs against the generator's ground truth.
prepare_cpu_side runs before CUDA is initialised (forked oracle workers); run_gpu_side afterwards.
"""
from __future__ import annotations

import ctypes as C
import os
import time

import numpy as np

from genomeassembler_dev_b200 import sharding, synth

SAMPLE_PER_RANK = 32
READ_LEN = 150
CFG5 = dict(truth=100_000_000, reads=100_000_000, contigs=100_000, lo=200, hi=2000, mut=0.05)
CFG5_MINI = dict(truth=1_000_000, reads=300_000, contigs=1000, lo=200, hi=2000, mut=0.05)
RTOL, KS_ATOL = 1e-9, 1e-12


def cfg5_host(seed, p):
    """cfg-5 shaped instance built with numpy (the mini one; also the CPU baseline's subsample)"""
    rng = np.random.default_rng(seed)
    L, N, Cn = p["truth"], p["reads"], p["contigs"]
    truth = synth.codes_to_ascii(synth.random_truth_codes(rng, L))
    starts = rng.integers(0, L - READ_LEN, size=N)
    reads = truth[starts[:, None] + np.arange(READ_LEN)[None, :]]
    cstart = np.sort(rng.integers(0, L - p["hi"] - 1, size=Cn))
    clen = rng.integers(p["lo"], p["hi"], size=Cn)
    contigs = []
    for a, b in zip(cstart, clen):
        c = truth[a:a + b].copy()
        if rng.random() < p["mut"]:
            q = int(rng.integers(0, b))
            c[q] = ord("ACGT"[("ACGT".index(chr(c[q])) + 1 + int(rng.integers(0, 3))) % 4])
        contigs.append(c.tobytes())
    return dict(truth=truth.tobytes(), reads=reads, contigs=contigs)


def prepare_cpu_side(bench, args, rank, world, procs):
    """oracle expectations for this rank's sample of every oracle-checkable config, and (one rank, N = 1) the CPU
    reference timed on a bounded sample of each config"""
    bench.load_checkers()
    plan = {"t_cpu_side_s": time.perf_counter()}
    cfg4 = synth.make_scaffold_set(1400, n_scaffolds=args.cfg4_scaffolds)
    mini = cfg5_host(1500, CFG5_MINI)
    for key, contigs, reads, truth in (("cfg4", cfg4.contigs, cfg4.reads, cfg4.truth), ("cfg5_mini", mini["contigs"], mini["reads"], mini["truth"])):
        lens = np.fromiter((len(c) for c in contigs), dtype=np.int64, count=len(contigs))
        parts = sharding.shard_contigs_lpt(lens, world)
        sample = bench.pick_sample(parts[rank], contigs, SAMPLE_PER_RANK, seed=4000 + rank)
        bench._POOL[key] = {"contigs": contigs, "read_list": [r.tobytes() for r in reads], "truth": truth}
        expected = bench.oracle_sample(key, sample, procs)
        plan[key] = {"contigs": contigs, "reads": np.ascontiguousarray(reads), "truth": truth, "parts": parts,
                     "sample": sample, "expected": expected, "seg": cfg4 if key == "cfg4" else None}
        if world == 1 and rank == 0 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            if key == "cfg4":
                rng = np.random.default_rng(7)
                pick = rng.choice(len(contigs), size=2 * cores, replace=False)
                groups = [pick[i::cores] for i in range(cores)]
                what = f"{2 * cores} of the {len(contigs)} scaffolds (2 per core) against all {len(reads)} reads"
            else:  # a stated subsample, scaled linearly (time is linear in reads x contig bases)
                sub = {"contigs": contigs, "read_list": bench._POOL[key]["read_list"][:10_000], "truth": truth}
                bench._POOL["cfg5_cpu"] = sub
                key = "cfg5_cpu"
                groups = [np.arange(len(contigs))[i::cores] for i in range(cores)]
                what = f"10 000 reads x {len(contigs)} contigs of the cfg5_mini instance (the survey's subsample), one process per core"
            cb = bench.cpu_rate(key, groups, cores)
            cb["sample"] = what + "; unmodified upstream calc_breakscore where built (edlib stubbed)"
            plan["cpu_" + ("cfg4" if key == "cfg4" else "cfg5")] = cb
        del bench._POOL[key]
    bench._POOL.pop("cfg5_mini", None)
    plan["t_cpu_side_s"] = time.perf_counter() - plan["t_cpu_side_s"]
    return plan


class Sharded:
    """one segment's contigs over the ranks of the process group; reads + truth are torch uint8 tensors on this rank's device"""

    def __init__(self, rig, tag, contigs, parts, d_reads, n_reads, d_truth):
        torch, B = rig.torch, rig.B
        self.rig, self.tag, self.contigs, self.parts = rig, tag, contigs, parts
        self.mine = parts[rig.rank]
        self.counts = [len(p) for p in parts]
        self.pad = max(max(self.counts), 1)
        self.n_reads, self.d_reads, self.d_truth = n_reads, d_reads, d_truth
        self.truth_len = int(d_truth.numel())
        self.ct, self.ct_off = B.flatten([contigs[i] for i in self.mine])
        self.d_ct = torch.from_numpy(self.ct).to(rig.dev)
        self.flags = B.WANT_KS | B.WANT_STARTPOS
        dev = rig.dev
        self.d_f64 = torch.zeros(5, self.pad, dtype=torch.float64, device=dev)
        self.d_i32 = torch.zeros(3, self.pad, dtype=torch.int32, device=dev)
        self.g_f64 = torch.zeros(rig.world, 5, self.pad, dtype=torch.float64, device=dev)
        self.g_i32 = torch.zeros(rig.world, 3, self.pad, dtype=torch.int32, device=dev)
        self.tr_off = np.array([0, self.truth_len], np.int64)
        self.srs = np.array([0, n_reads], np.int64)
        self.scs = np.array([0, len(self.mine)], np.int64)
        self.res = self.result_struct(self.d_f64, self.d_i32)
        self.bases = float(sum(len(c) for c in contigs))

    def result_struct(self, f64, i32):
        r = self.rig.B._Result()
        r.sequence_len, r.kmer_breaks, r.path_prob_dist_startpos = [i32[i].data_ptr() for i in range(3)]
        (r.bp_score, r.bp_score_norm_by_break_freqs, r.bp_score_norm_by_len, r.ks_stat_prob_dist,
         r.ks_stat_path_freq) = [f64[i].data_ptr() for i in range(5)]
        return r

    def batch_struct(self, rc, cc, tc, n_contigs=None, ct_off=None):
        B = self.rig.B
        scs = self.scs if n_contigs is None else np.array([0, n_contigs], np.int64)
        self._keep = scs
        return B._Batch(1, self.n_reads, int(scs[1]), rc, None, READ_LEN, cc, (self.ct_off if ct_off is None else ct_off).ctypes.data, tc,
                        self.tr_off.ctypes.data, self.srs.ctypes.data, scs.ctypes.data)

    def gather(self):
        if self.rig.world > 1:
            d = self.rig.dist
            d.all_gather_into_tensor(self.g_f64.view(-1), self.d_f64.view(-1))
            d.all_gather_into_tensor(self.g_i32.view(-1), self.d_i32.view(-1))
        else:
            self.g_f64[0].copy_(self.d_f64)
            self.g_i32[0].copy_(self.d_i32)

    def step_device(self):
        B, sc = self.rig.B, self.rig.sc
        if len(self.mine):
            sc.score_batch_raw(self.batch_struct(self.d_reads.data_ptr(), self.d_ct.data_ptr(), self.d_truth.data_ptr()), self.res, 8,
                               self.flags | B.DEVICE_CHARS | B.DEVICE_RESULT)
        self.gather()

    def table_in_input_order(self):
        """(f64 [C,5], i32 [C,3]) of the gathered records, rows in input contig order (device tensors)"""
        torch = self.rig.torch
        order = np.concatenate(self.parts) if len(self.contigs) else np.zeros(0, np.int64)
        f = torch.cat([self.g_f64[r, :, :self.counts[r]] for r in range(self.rig.world)], dim=1).t().contiguous()
        i = torch.cat([self.g_i32[r, :, :self.counts[r]] for r in range(self.rig.world)], dim=1).t().contiguous()
        idx = torch.from_numpy(order).to(self.rig.dev)
        tf, ti = torch.empty_like(f), torch.empty_like(i)
        tf[idx] = f
        ti[idx] = i
        return tf, ti

    def check_sample(self, sample, expected):
        """this rank's records of the sampled contigs against the CPU oracle's: ints bit-exact, fp64 1e-9 relative"""
        if not len(sample):
            return True, 0
        loc = np.searchsorted(self.mine, sample)
        assert np.array_equal(self.mine[loc], sample)
        f = self.d_f64.cpu().numpy()[:, loc]
        i = self.d_i32.cpu().numpy()[:, loc]
        ok = True
        for j, k in enumerate(sharding.RECORD_I32):
            ok = ok and bool(np.array_equal(i[j], expected[k]))
        for j, k in enumerate(sharding.RECORD_F64):
            atol = KS_ATOL if k.startswith("ks_") else 0.0
            ok = ok and bool(np.allclose(f[j], expected[k], rtol=RTOL, atol=atol, equal_nan=True))
        return ok, len(sample)


def measure(bench, rig, sh, steps, warmup, sample=None, expected=None, host_reads=None, host_truth=None, timing=True,
            one_gpu_check=True, multi_ctx=False):
    torch, dist, B, sc = rig.torch, rig.dist, rig.B, rig.sc
    world, rank = rig.world, rig.rank
    out = {"contigs": len(sh.contigs), "contig_bases": int(sh.bases), "reads": int(sh.n_reads), "read_len": READ_LEN,
           "contigs_per_rank": sh.counts, "scaling": "strong", "unit": bench.UNIT,
           "partitioning": f"contigs of ONE segment dealt out longest-first over {world} rank(s), reads and truth replicated, "
                           f"8-column records all-gathered over NCCL into input order"}
    pair = float(sh.n_reads) * sh.bases
    dev_ms = rig.time_device(sh.step_device, steps if timing else 1, warmup if timing else 1)
    if timing:
        out["ms_per_step"] = dev_ms
        out["value"] = pair / 1e9 / (dev_ms / 1e3)
        out["reads_scored_per_s"] = sh.n_reads / (dev_ms / 1e3)
        sc.enable_timing(True)  # one more (untimed) step with the stage events on
        sh.step_device()
        rig.sync()
        out["stage_ms_rank0_extra_step"] = {k: v for k, v in sc.last_timings().items() if v > 0}
        sc.enable_timing(False)
    # ---- parity: sampled oracle diff on every rank ----
    if sample is not None:
        ok, n = sh.check_sample(sample, expected)
        tot = rig.sum_over_ranks([0.0 if ok else 1.0, float(n)])
        out["oracle_sample"] = {"contigs_checked": int(tot[1]), "per_rank": SAMPLE_PER_RANK, "all_match": tot[0] == 0.0,
                                "checker": "oracle/ C restatement of lib/BreakageScorer.cpp:185-353 + ks.test statistic, ALL reads of the segment; "
                                           "integer columns bit-exact, fp64 columns rtol 1e-9 (KS atol 1e-12)"}
    # ---- parity: the gathered table == ONE GPU scoring every contig (rank 0 alone; the others wait) ----
    tf, ti = sh.table_in_input_order()
    if one_gpu_check:
        rig.sync()
        rig.barrier()
        if rank == 0 and world > 1:
            ct, ct_off = B.flatten(sh.contigs)
            d_ct = torch.from_numpy(ct).to(rig.dev)
            C = len(sh.contigs)
            f1 = torch.zeros(5, C, dtype=torch.float64, device=rig.dev)
            i1 = torch.zeros(3, C, dtype=torch.int32, device=rig.dev)
            sc.score_batch_raw(sh.batch_struct(sh.d_reads.data_ptr(), d_ct.data_ptr(), sh.d_truth.data_ptr(), n_contigs=C, ct_off=ct_off),
                               sh.result_struct(f1, i1), 8, sh.flags | B.DEVICE_CHARS | B.DEVICE_RESULT)
            rig.sync()
            same = bool(torch.equal(f1.t().nan_to_num(nan=-1.0), tf.nan_to_num(nan=-1.0)) and torch.equal(i1.t(), ti))
            out["identical_to_one_gpu_scoring_all_contigs"] = same
            del d_ct, f1, i1
        rig.barrier()
    out["checksum"] = {"kmer_breaks_sum": int(ti[:, 1].sum().item()), "startpos_minus1": int((ti[:, 2] == -1).sum().item()),
                       "bp_score_sum": float(tf[:, 0].sum().item())}
    if not timing:
        return out
    # ---- host-to-host, two ways of replicating the reads ----
    if host_reads is not None:
        h_ct = rig.pinned(sh.ct)
        h_out_f = torch.zeros_like(sh.g_f64, device="cpu").pin_memory()
        h_out_i = torch.zeros_like(sh.g_i32, device="cpu").pin_memory()

        def results_home():
            sh.gather()
            if rank == 0:
                h_out_f.copy_(sh.g_f64, non_blocking=True)
                h_out_i.copy_(sh.g_i32, non_blocking=True)

        def step_each():  # every rank copies ALL the reads in (the library's own staged H2D)
            if len(sh.mine):
                sc.score_batch_raw(sh.batch_struct(host_reads.data_ptr(), h_ct.data_ptr(), host_truth.data_ptr()), sh.res, 8,
                                   sh.flags | B.DEVICE_RESULT)
            results_home()

        n_e2e = max(2, min(steps, 3))
        each_ms = rig.time_wall(step_each, n_e2e, 1)
        ok_each = bool(torch.equal(h_out_f.nan_to_num(nan=-1.0), sh.g_f64.cpu().nan_to_num(nan=-1.0))) if rank == 0 else True
        # every rank copies 1/N of the reads in, NCCL all-gathers the rest over NVLink
        chunk = (sh.n_reads + world - 1) // world
        d_full = torch.empty(world * chunk * READ_LEN, dtype=torch.uint8, device=rig.dev)
        r0, r1 = min(rank * chunk, sh.n_reads), min((rank + 1) * chunk, sh.n_reads)
        my_slice = d_full[rank * chunk * READ_LEN:(rank + 1) * chunk * READ_LEN]
        d_ct2, d_tr2 = torch.empty_like(sh.d_ct), torch.empty_like(sh.d_truth)

        def step_shard():
            if r1 > r0:
                my_slice[:(r1 - r0) * READ_LEN].copy_(host_reads[r0 * READ_LEN:r1 * READ_LEN], non_blocking=True)
            d_ct2.copy_(h_ct, non_blocking=True)
            d_tr2.copy_(host_truth, non_blocking=True)
            if world > 1:
                dist.all_gather_into_tensor(d_full, my_slice)
            if len(sh.mine):
                sc.score_batch_raw(sh.batch_struct(d_full.data_ptr(), d_ct2.data_ptr(), d_tr2.data_ptr()), sh.res, 8,
                                   sh.flags | B.DEVICE_CHARS | B.DEVICE_RESULT)
            results_home()

        shard_ms = rig.time_wall(step_shard, n_e2e, 1)
        ok_shard = bool(torch.equal(h_out_f.nan_to_num(nan=-1.0), sh.g_f64.cpu().nan_to_num(nan=-1.0))) if rank == 0 else True
        h2d_each = int(host_reads.numel() + h_ct.numel() + host_truth.numel())
        h2d_shard = int((r1 - r0) * READ_LEN + h_ct.numel() + host_truth.numel())
        out["e2e"] = {"value": pair / 1e9 / (min(each_ms, shard_ms) / 1e3), "unit": bench.UNIT, "ms_per_step": min(each_ms, shard_ms),
                      "d2h_bytes_per_step": int(h_out_f.numel() * 8 + h_out_i.numel() * 4),
                      "h2d_bytes_per_step": h2d_each if each_ms <= shard_ms else h2d_shard,
                      "same_records_as_device_resident": ok_each and ok_shard}
        out["read_replication"] = {
            "every_rank_copies_all_reads": {"ms_per_step": each_ms, "h2d_bytes_per_rank": h2d_each},
            "each_rank_copies_its_share_then_nccl_all_gather": {"ms_per_step": shard_ms, "h2d_bytes_per_rank": h2d_shard,
                                                                "nvlink_bytes_received_per_rank": int((world - 1) * chunk * READ_LEN)},
            "what": "host ASCII -> records on rank 0's host, contigs sharded either way; the reads reach every GPU over PCIe (xN) or "
                    "once over PCIe + NVLink"}
        del d_full, d_ct2, d_tr2
    if multi_ctx and world > 1 and host_reads is not None:
        one = measure_one_process(bench, rig, sh.contigs, sh.tr_off, sh.srs, sh.flags, host_reads, host_truth, tf, ti, pair)
        if one is not None:
            out["one_process_bs_score_multi"] = one
    return out


def measure_compositional(bench, rig, sh, seg, steps, warmup, sample, expected, host_reads, host_truth):
    """cfg-4 scored from the PARTS of its scaffolds (bs_score_scaffolds): the reads are placed once per base
    contig, every scaffold takes the minimum over its parts + junction-window probes and is scored from those positions.
    Same sharding as the rescan above (this rank's scaffolds, all base contigs), same record buffers, same gather; the
    gathered table is compared with the rescan's."""
    torch, B, sc = rig.torch, rig.B, rig.sc
    world, rank = rig.world, rig.rank
    tf0, ti0 = sh.table_in_input_order()
    tf0, ti0 = tf0.clone(), ti0.clone()
    # this rank's share of the scaffolds as a set of its own
    ps = seg.part_start
    counts = np.diff(ps)[sh.mine]
    my_start = np.zeros(len(sh.mine) + 1, np.int64)
    np.cumsum(counts, out=my_start[1:])
    my_base = np.concatenate([seg.part_base[ps[i]:ps[i + 1]] for i in sh.mine]) if len(sh.mine) else np.zeros(0, np.int32)
    sset = B.ScaffoldSet(seg.base_contigs, my_start, my_base, np.zeros(len(my_base), np.int32))
    st = sset.c_struct()
    lib, ctx = sc._lib, sc._ctx

    def call(rc, tc, flags):
        if len(sh.mine):
            sc._check(lib.bs_score_scaffolds(ctx, C.byref(st), rc, None, sh.n_reads, READ_LEN, tc, sh.truth_len, 8, flags, C.byref(sh.res)))

    def step_device():
        call(sh.d_reads.data_ptr(), sh.d_truth.data_ptr(), sh.flags | B.DEVICE_CHARS | B.DEVICE_RESULT)
        sh.gather()

    dev_ms = rig.time_device(step_device, steps, warmup)
    pair = float(sh.n_reads) * sh.bases
    out = {"ms_per_step": dev_ms, "value": pair / 1e9 / (dev_ms / 1e3), "unit": bench.UNIT, "scaffolds_per_rank": sh.counts,
           "base_contigs": len(seg.base_contigs), "parts": int(ps[-1]),
           "what": "bs_score_scaffolds: reads placed once per base contig (k_place_index), scaffolds from their parts + junction-window "
                   "probes and scored from those positions (k_place_compose); scaffold texts composed on the device (KS-A, startpos "
                   "still scan them); value = the same read x scaffold-base pairs as the rescan, per second"}
    sc.enable_timing(True)
    step_device()
    rig.sync()
    out["stage_ms_rank0_extra_step"] = {k: v for k, v in sc.last_timings().items() if v > 0}
    sc.enable_timing(False)
    ok, n = sh.check_sample(sample, expected)
    tot = rig.sum_over_ranks([0.0 if ok else 1.0, float(n)])
    out["oracle_sample"] = {"contigs_checked": int(tot[1]), "all_match": tot[0] == 0.0}
    tf, ti = sh.table_in_input_order()
    f0, f1 = tf0.nan_to_num(nan=-1.0), tf.nan_to_num(nan=-1.0)
    rel = ((f1 - f0).abs() / f0.abs().clamp_min(1e-300)).max().item() if f0.numel() else 0.0
    out["vs_rescan"] = {"integer_columns_identical": bool(torch.equal(ti, ti0)), "ks_prob_dist_identical": bool(torch.equal(f1[:, 3], f0[:, 3])),
                        "max_rel_diff_fp64_columns": rel, "within_1e-9": rel <= 1e-9}
    if host_reads is not None:
        h_out_f = torch.zeros_like(sh.g_f64, device="cpu").pin_memory()
        h_out_i = torch.zeros_like(sh.g_i32, device="cpu").pin_memory()

        def step_host():
            call(host_reads.data_ptr(), host_truth.data_ptr(), sh.flags | B.DEVICE_RESULT)
            sh.gather()
            if rank == 0:
                h_out_f.copy_(sh.g_f64, non_blocking=True)
                h_out_i.copy_(sh.g_i32, non_blocking=True)

        ms = rig.time_wall(step_host, max(2, min(steps, 3)), 1)
        out["e2e"] = {"ms_per_step": ms, "value": pair / 1e9 / (ms / 1e3), "unit": bench.UNIT,
                      "h2d_bytes_per_step": int(host_reads.numel() + host_truth.numel() + sset.base_chars.nbytes + 12 * len(my_base)),
                      "d2h_bytes_per_step": int(h_out_f.numel() * 8 + h_out_i.numel() * 4),
                      "same_records_as_device_resident": bool(torch.equal(h_out_i, sh.g_i32.cpu())) if rank == 0 else True}
    return out


def measure_one_process(bench, rig, contigs, tr_off, srs, flags, host_reads, host_truth, tf, ti, pair):
    """ONE process driving all GPUs (upstream's R driver is one process): bs_score_multi from rank 0 while the other ranks
    wait; host buffers in, host arrays out in input order, no collective.  A read set of 64 MB or more crosses PCIe once
    and is replicated by peer copies over NVLink (BS_MULTI_P2P_MB); timed both ways when the set is that large."""
    B, sc, world, rank = rig.B, rig.sc, rig.world, rig.rank
    rig.sync()
    rig.barrier()
    res1 = None
    if rank == 0:
        others = []
        try:
            others = [B.BreakageScorer(d) for d in range(1, world)]
            for o in others:
                o.set_table(rig.kmers, rig.prob)
            ct, ct_off = B.flatten(contigs)
            rd = host_reads.numpy()
            tr = host_truth.numpy()
            call = lambda: sc.score_batch(rd, None, READ_LEN, ct, ct_off, tr, tr_off, srs, [0, len(contigs)], flags=flags, group=others)  # noqa: E731
            res1 = {"contexts": world,
                    "what": "ONE process, one context per GPU, a host thread each (bs_score_multi): host buffers in, host arrays out in "
                            "input order, no collective; a read set of 64 MB or more crosses PCIe once and is replicated by peer copies"}
            big = rd.nbytes >= (64 << 20)
            for name, env in ((("reads_by_peer_copies", None), ("reads_over_pcie_per_gpu", "-1")) if big else (("reads_over_pcie_per_gpu", None),)):
                if env is None:
                    os.environ.pop("BS_MULTI_P2P_MB", None)
                else:
                    os.environ["BS_MULTI_P2P_MB"] = env
                try:
                    call()
                    t0 = time.perf_counter()
                    got = call()
                    ms = 1e3 * (time.perf_counter() - t0)
                    same = bool(np.array_equal(got["kmer_breaks"], ti[:, 1].cpu().numpy()) and
                                np.array_equal(got["bp_score"], tf[:, 0].cpu().numpy()))
                    res1[name] = {"ms_per_call": ms, "value": pair / 1e9 / (ms / 1e3), "unit": bench.UNIT, "same_records": same}
                except B.BreakscoreError as e:  # e.g. device memory: the other ranks' buffers share these GPUs
                    res1[name] = {"failed": str(e)[:200]}
            os.environ.pop("BS_MULTI_P2P_MB", None)
        finally:
            for o in others:
                o.close()
    rig.barrier()
    return res1


def measure_read_sharded(bench, rig, sh, steps, warmup, timing=True):
    """The other split of the same segment (sharding.score_reads_sharded on device-resident inputs): every rank packs,
    indexes and places 1/N of the READS against ALL contigs, the position weights are all-reduced over NCCL (the path's
    one real data exchange: int32, sum(L_c) + C entries), every rank scores a contiguous range of the contigs from the
    sum, the records are all-gathered.  Bit-identical to the contig-sharded table (additivity of the break counts)."""
    torch, dist, B, sc = rig.torch, rig.dist, rig.B, rig.sc
    world, rank, dev = rig.world, rig.rank, rig.dev
    lens = np.fromiter((len(c) for c in sh.contigs), dtype=np.int64, count=len(sh.contigs))
    ct, ct_off = B.flatten(sh.contigs)
    d_ct = torch.from_numpy(ct).to(dev)
    C = len(sh.contigs)
    bounds = sharding.balanced_ranges(lens, world)
    c0, c1 = bounds[rank], bounds[rank + 1]
    counts = [bounds[r + 1] - bounds[r] for r in range(world)]
    pad = max(max(counts), 1)
    n0, n1 = sh.n_reads * rank // world, sh.n_reads * (rank + 1) // world
    w = torch.zeros(int(ct_off[-1]) + C + 1, dtype=torch.int32, device=dev)
    tot = torch.zeros(C + 1, dtype=torch.int32, device=dev)
    f64 = torch.zeros(5, pad, dtype=torch.float64, device=dev)
    i32 = torch.zeros(3, pad, dtype=torch.int32, device=dev)
    g_f64 = torch.zeros(world, 5, pad, dtype=torch.float64, device=dev)
    g_i32 = torch.zeros(world, 3, pad, dtype=torch.int32, device=dev)
    tr_off = sh.tr_off
    srs1, scs1 = np.array([0, n1 - n0], np.int64), np.array([0, C], np.int64)
    srs2, scs2 = np.array([0, 0], np.int64), np.array([0, c1 - c0], np.int64)
    off2 = np.ascontiguousarray(ct_off[c0:c1 + 1] - ct_off[c0])
    b1 = B._Batch(1, n1 - n0, C, sh.d_reads.data_ptr() + n0 * READ_LEN, None, READ_LEN, d_ct.data_ptr(), ct_off.ctypes.data,
                  sh.d_truth.data_ptr(), tr_off.ctypes.data, srs1.ctypes.data, scs1.ctypes.data)
    r1 = B._Result()
    r1.weights, r1.weights_total = w.data_ptr(), tot.data_ptr()
    b2 = B._Batch(1, 0, c1 - c0, None, None, READ_LEN, d_ct.data_ptr() + int(ct_off[c0]), off2.ctypes.data,
                  sh.d_truth.data_ptr(), tr_off.ctypes.data, srs2.ctypes.data, scs2.ctypes.data)
    r2 = sh.result_struct(f64, i32)
    r2.weights, r2.weights_total = w.data_ptr() + 4 * (int(ct_off[c0]) + c0), tot.data_ptr() + 4 * c0
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]

    def step():
        ev[0].record()
        sc.score_batch_raw(b1, r1, 8, B.DEVICE_CHARS | B.WEIGHTS_OUT)
        ev[1].record()
        if world > 1:
            dist.all_reduce(w)
            dist.all_reduce(tot)
        ev[2].record()
        if c1 > c0:
            sc.score_batch_raw(b2, r2, 8, sh.flags | B.DEVICE_CHARS | B.DEVICE_RESULT | B.WEIGHTS_IN)
        if world > 1:
            dist.all_gather_into_tensor(g_f64.view(-1), f64.view(-1))
            dist.all_gather_into_tensor(g_i32.view(-1), i32.view(-1))
        else:
            g_f64[0].copy_(f64)
            g_i32[0].copy_(i32)
        ev[3].record()

    ms = rig.time_device(step, steps if timing else 1, warmup if timing else 1)
    rig.sync()
    phases = [rig.max_over_ranks(ev[i].elapsed_time(ev[i + 1])) for i in range(3)]
    tf = torch.cat([g_f64[r, :, :counts[r]] for r in range(world)], dim=1).t().contiguous()
    ti = torch.cat([g_i32[r, :, :counts[r]] for r in range(world)], dim=1).t().contiguous()
    rf, ri = sh.table_in_input_order()
    out = {"same_table_as_contig_sharded": bool(torch.equal(tf.nan_to_num(nan=-1.0), rf.nan_to_num(nan=-1.0)) and torch.equal(ti, ri)),
           "reads_per_rank": int(n1 - n0), "contigs_scored_per_rank": counts,
           "all_reduce_bytes": int(w.numel() * 4 + tot.numel() * 4),
           "what": "reads sharded: phase 1 = pack + index + place this rank's reads in ALL contigs (BS_WEIGHTS_OUT); all-reduce of the "
                   "int32 position weights over NCCL; phase 2 = score a contiguous contig range from the sum (BS_WEIGHTS_IN); records "
                   "all-gathered"}
    if timing:
        pair = float(sh.n_reads) * sh.bases
        out.update({"ms_per_step": ms, "value": pair / 1e9 / (ms / 1e3), "unit": bench.UNIT,
                    "last_step_ms_max_over_ranks": {"phase1_pack_index_place": phases[0], "all_reduce_weights": phases[1],
                                                    "phase2_score_and_gather": phases[2]}})
    del w, tot, d_ct
    return out


def cfg5_device(rig, p, seed=1500):
    """cfg-5 on the device: truth, reads (uniform starts), contigs (substrings, a fraction with one substitution) and the
    generator's ground truth for kmer_breaks / startpos of every contig"""
    torch = rig.torch
    dev = rig.dev
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    L, N, Cn = p["truth"], p["reads"], p["contigs"]
    lut = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device=dev)
    truth = lut[torch.randint(0, 4, (L,), generator=g, device=dev, dtype=torch.int32)]
    starts = torch.randint(0, L - READ_LEN, (N,), generator=g, device=dev, dtype=torch.int32)
    reads = torch.empty(N * READ_LEN, dtype=torch.uint8, device=dev)
    ar = torch.arange(READ_LEN, device=dev, dtype=torch.int32)
    step = 2_000_000
    for a in range(0, N, step):
        b = min(N, a + step)
        reads[a * READ_LEN:b * READ_LEN] = truth[(starts[a:b, None] + ar[None, :])].reshape(-1)
    cstart = torch.sort(torch.randint(0, L - p["hi"] - 1, (Cn,), generator=g, device=dev, dtype=torch.int64)).values
    clen = torch.randint(p["lo"], p["hi"], (Cn,), generator=g, device=dev, dtype=torch.int64)
    mut = torch.rand(Cn, generator=g, device=dev) < p["mut"]
    mpos = (torch.rand(Cn, generator=g, device=dev) * clen).long().clamp(max=clen - 1)
    mdelta = torch.randint(1, 4, (Cn,), generator=g, device=dev)
    # ground truth: reads whose start lies in [a, a + len - r] and (mutated contigs) that do not cover the substituted base;
    # exact because a random 100 Mb truth holds no repeated 150-mer
    ss = torch.sort(starts.long()).values
    cnt = lambda lo, hi: (torch.searchsorted(ss, hi, right=True) - torch.searchsorted(ss, lo, right=False)).clamp(min=0)  # noqa: E731
    a0, a1 = cstart, cstart + clen - READ_LEN
    plain = cnt(a0, a1) * (a1 >= a0)
    pm = cstart + mpos
    left_hi, right_lo = torch.minimum(a1, pm - READ_LEN), torch.maximum(a0, pm + 1)
    split = cnt(a0, left_hi) * (left_hi >= a0) + cnt(right_lo, a1) * (a1 >= right_lo)
    breaks = torch.where(mut, split, plain).int()
    startpos = torch.where(mut, torch.full_like(cstart, -1), cstart).int() * (breaks > 0).int()
    # contigs as host bytes (they are sharded on the host)
    h_truth = truth.cpu().numpy()
    cs, cl, mu, mp, md = (x.cpu().numpy() for x in (cstart, clen, mut, mpos, mdelta))
    acgt = b"ACGT"
    contigs = []
    for i in range(Cn):
        c = h_truth[cs[i]:cs[i] + cl[i]]
        if mu[i]:
            c = c.copy()
            c[mp[i]] = acgt[(acgt.index(int(c[mp[i]])) + int(md[i])) % 4]
        contigs.append(c.tobytes())
    del ss, starts
    return dict(d_truth=truth, d_reads=reads, contigs=contigs, breaks=breaks, startpos=startpos, n_reads=N)


def run_gpu_side(bench, rig, plan):
    torch = rig.torch
    args, world, rank, dev = rig.args, rig.world, rig.rank, rig.dev
    out = {"cpu_side_seconds": plan["t_cpu_side_s"]}
    steps, warmup = args.steps, args.warmup

    # ---- cfg4 and the oracle-checkable cfg5 instance ----
    for key in ("cfg4", "cfg5_mini"):
        d = plan[key]
        reads = d["reads"].reshape(-1)
        d_reads = torch.from_numpy(reads).to(dev)
        d_truth = torch.from_numpy(np.frombuffer(d["truth"], np.uint8).copy()).to(dev)
        sh = Sharded(rig, key, d["contigs"], d["parts"], d_reads, len(d["reads"]), d_truth)
        timing = key == "cfg4"
        res = measure(bench, rig, sh, steps, warmup, sample=d["sample"], expected=d["expected"],
                      host_reads=rig.pinned(reads) if timing else None,
                      host_truth=rig.pinned(np.frombuffer(d["truth"], np.uint8)) if timing else None,
                      timing=timing, multi_ctx=timing)
        if key == "cfg4" and d.get("seg") is not None:
            res["compositional"] = measure_compositional(bench, rig, sh, d["seg"], steps, warmup, d["sample"], d["expected"],
                                                         rig.pinned(reads), rig.pinned(np.frombuffer(d["truth"], np.uint8)))
        if key == "cfg4":
            res["workload"] = (f"cfg4: {len(d['contigs'])} candidate scaffolds of 10-50 kb (concatenations of 16 base contigs) of one 50 kb "
                               f"segment, {len(d['reads'])} reads of 150 bp (30x), outputs scores+kmer_breaks+startpos+KS")
            if "cpu_cfg4" in plan:
                res["cpu_baseline"] = plan["cpu_cfg4"]
        else:
            res["workload"] = "cfg5 shape at 1 Mb truth / 3e5 reads / 1000 contigs (5 % mutated): parity only, the CPU oracle sees every read"
            res["read_sharded"] = measure_read_sharded(bench, rig, sh, steps, warmup, timing=False)
        out[key] = res
        del sh, d_reads, d_truth
        torch.cuda.empty_cache()

    # ---- cfg5 at (a fraction of) full size ----
    p = dict(CFG5)
    if args.cfg5_scale != 1.0:
        for k in ("truth", "reads", "contigs"):
            p[k] = max(1000, int(p[k] * args.cfg5_scale))
    t0 = time.perf_counter()
    g = cfg5_device(rig, p)
    rig.sync()
    gen_s = time.perf_counter() - t0
    lens = np.fromiter((len(c) for c in g["contigs"]), dtype=np.int64, count=len(g["contigs"]))
    parts = sharding.shard_contigs_lpt(lens, world)
    sh = Sharded(rig, "cfg5", g["contigs"], parts, g["d_reads"], g["n_reads"], g["d_truth"])
    # pinned host copy of the reads for the host-to-host numbers, if the box has the memory for N of them
    need = g["d_reads"].numel() * 1.15 * world
    avail = 0.0
    try:
        with open("/proc/meminfo") as fh:
            for line in fh:
                if line.startswith("MemAvailable"):
                    avail = float(line.split()[1]) * 1024
    except OSError:
        pass
    h_reads = h_truth = None
    enough = rig.sum_over_ranks([1.0 if avail > need + 16e9 else 0.0])[0] == world
    if enough:
        h_reads = torch.empty(g["d_reads"].numel(), dtype=torch.uint8, pin_memory=True)
        h_reads.copy_(g["d_reads"])
        h_truth = torch.empty(g["d_truth"].numel(), dtype=torch.uint8, pin_memory=True)
        h_truth.copy_(g["d_truth"])
    res = measure(bench, rig, sh, steps, warmup, host_reads=h_reads, host_truth=h_truth, timing=True)
    res["read_sharded"] = measure_read_sharded(bench, rig, sh, steps, warmup)
    if not enough:
        res["e2e"] = {"skipped": f"host MemAvailable {avail / 1e9:.0f} GB < {need / 1e9:.0f} GB needed to pin the reads on {world} rank(s)"}
    # ground truth of the generator for ALL contigs (kmer_breaks, startpos), on the gathered table
    tf, ti = sh.table_in_input_order()
    res["ground_truth_all_contigs"] = {
        "kmer_breaks_equal": bool(torch.equal(ti[:, 1], g["breaks"])), "startpos_equal": bool(torch.equal(ti[:, 2], g["startpos"])),
        "sequence_len_equal": bool(torch.equal(ti[:, 0].cpu(), torch.from_numpy(lens).int())),
        "contigs": len(lens), "mutated": int((g["startpos"] == -1).sum().item()),
        "what": "reads placed per contig and contig-in-truth offset of every contig against the generator: reads whose start lies "
                "inside the contig (and, for the contigs with a substitution, that do not cover it); exact because a random truth "
                "of this size holds no repeated 150-mer"}
    res["workload"] = (f"cfg5: {p['truth'] / 1e6:.0f} Mb synthetic truth, {p['reads']:.3g} reads of 150 bp (uniform starts), {p['contigs']} contigs "
                       f"of {p['lo']}-{p['hi']} bp ({100 * p['mut']:.0f} % with one substitution), generated on the device in {gen_s:.1f} s; "
                       f"scale {args.cfg5_scale:g} of BASELINE.json configs[4]")
    if "cpu_cfg5" in plan:
        res["cpu_baseline"] = plan["cpu_cfg5"]
    out["cfg5"] = res
    # the one-process path last, with this rank's device tensors gone (rank 0's extra contexts need room on every GPU)
    contigs, tr_off, srs, flags, pair = sh.contigs, sh.tr_off, sh.srs, sh.flags, float(sh.n_reads) * sh.bases
    tf, ti = tf.clone(), ti.clone()
    del sh, g
    torch.cuda.empty_cache()
    if world > 1 and h_reads is not None:
        one = measure_one_process(bench, rig, contigs, tr_off, srs, flags, h_reads, h_truth, tf, ti, pair)
        if one is not None:
            res["one_process_bs_score_multi"] = one
    del h_reads, h_truth
    return out
