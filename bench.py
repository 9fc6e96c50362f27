#!/usr/bin/env python
"""Benchmark of the breakage-scoring hot path (BASELINE.json metric: reads scored/s and
read x contig Gbp compared/s; % of HBM peak).

A "step" is one pass of the hot path over one batch: the cfg-2 study of BASELINE.json
(`--segments` synthetic 50 kb segments, 150 bp reads at 30x, velvet-style contig sets) scored by ONE
bs_score_batch call per rank with the upstream default outputs (scores, kmer_breaks, startpos,
path_prob_dist, KS statistics).  Ranks own disjoint studies (weak scaling: segments are independent
units, no data-path collective; per-contig score records are gathered to rank 0 over NCCL).

  value     whole-job Gbp/s, inputs already resident in HBM (device pointers through the C-ABI)
  e2e       the same metric through the C-ABI with HOST buffers: H2D of the ASCII inputs and D2H of
            every result array inside the timed region
  roofline  placement kernel: algorithmic bytes ((8*W_r+4) B per (unique read, contig)
            pair) / its CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline / --impl reference: the UNMODIFIED upstream calc_breakscore (oracle/_ref, edit distance
            stubbed: off the scored path) on the box's host cores, one process per core over
            disjoint segments of the same workload (bounded sample)
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from genomeassembler_dev_b200 import synth, tables  # noqa: E402

LAST_PER_CORE = None
METRIC = "read_x_contig_Gbp_compared_per_s"
UNIT = "Gbp/s"
SEED = 1234
LENGTH, READ_LEN, COVERAGE = 50_000, 150, 30.0


def workload_name(n_segments):
    return (f"cfg2: {n_segments} synthetic 50 kb segments per GPU, 150 bp reads 30x, velvet-style contig sets "
            f"(5-60 contigs/segment, 10% mutated), kmer=8, real breakage table")


# ------------------------------------------------------------------------------------------
# CPU reference arm (test/bench infrastructure: oracle/_ref = unmodified upstream code)
# ------------------------------------------------------------------------------------------

def _ref_task(args):
    seeds, use_ref = args
    from oracle import loader as O
    kmers = tables.all_kmer_strings()
    prob = tables.normalised(tables.load_raw())
    p8 = tables.sub_table(prob, 8)
    segs = []
    for s in seeds:
        rng = np.random.default_rng(s)
        nct = int(rng.integers(5, 61))
        segs.append(synth.make_segment(s, LENGTH, READ_LEN, COVERAGE, nct, p8))
    reads = [sg.read_list for sg in segs]
    t0 = time.perf_counter()
    pair = 0.0
    nreads = 0
    for sg, rl in zip(segs, reads):
        if use_ref:
            O.ref_calc_breakscore(sg.contigs, rl, sg.truth, 8, kmers, prob, edit_distance=False)
        else:
            O.oracle_calc_breakscore(sg.contigs, rl, sg.truth, 8, kmers, prob, want_ks=False)
        pair += len(rl) * float(sum(len(c) for c in sg.contigs))
        nreads += len(rl)
    return time.perf_counter() - t0, pair, nreads


def run_cpu_reference(n_segments_sample, cores, seed0=SEED):
    """one process per core, disjoint segments; returns (Gbp/s, reads/s, seconds, kind)"""
    import multiprocessing as mp
    from oracle import loader as O
    O.build()
    use_ref = O.have_ref()
    seeds = [seed0 + i for i in range(n_segments_sample)]
    chunks = [seeds[i::cores] for i in range(cores) if seeds[i::cores]]
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(len(chunks)) as pool:
        res = pool.map(_ref_task, [(c, use_ref) for c in chunks])
    wall_all = time.perf_counter() - t0
    t = max(r[0] for r in res)  # scoring time of the slowest process (generation excluded)
    pair = sum(r[1] for r in res)
    nreads = sum(r[2] for r in res)
    global LAST_PER_CORE
    LAST_PER_CORE = float(np.mean([r[1] / 1e9 / r[0] for r in res]))  # what one core does (the reference is single-threaded)
    return pair / 1e9 / t, nreads / t, t, ("reference" if use_ref else "port"), wall_all


def reference_main(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    per_step = max(cores, 8)
    for _ in range(args.warmup):
        run_cpu_reference(min(per_step, cores), cores)
    tot_pair, tot_reads, tot_t = 0.0, 0.0, 0.0
    kind = "reference"
    for k in range(args.steps):
        gbps, rps, t, kind, _ = run_cpu_reference(per_step, cores, seed0=SEED + k * per_step)
        tot_pair += gbps * t
        tot_reads += rps * t
        tot_t += t
    value = tot_pair / tot_t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8 compare / int32 count / f64 sums",
        "data": "synthetic", "reads_scored_per_s": tot_reads / tot_t,
        "config": {"workload": workload_name(args.segments), "sample_per_step": f"{per_step} segments of that workload"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "per_core_value": LAST_PER_CORE,
                         "sample": f"{per_step} segments per step, one process per core, unmodified upstream "
                                   f"calc_breakscore (edlib stubbed)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------

class ClockSampler:
    """SM clock and throttle reasons sampled through NVML every few ms DURING the timed region
    (nvidia-smi -lms cannot resolve a timed region of tens of milliseconds)."""

    def __init__(self, index, period_s=0.002):
        self.index, self.period = index, period_s
        self.mhz, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self.thread = None
        self.err = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.index
            if vis:
                try:
                    idx = int(vis.split(",")[self.index])
                except ValueError:
                    idx = self.index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # no NVML: report it, do not guess
            self.err = repr(e)
            return
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        nv = self.nv
        names = {nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksEventReasonHwPowerBrakeSlowdown: "hw_power_brake_slowdown"}
        while not self._stop.is_set():
            try:
                self.mhz.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception as e:
                self.err = repr(e)
                return
            time.sleep(self.period)

    def mark(self):
        self.first = len(self.mhz)

    def stop(self):
        self._stop.set()
        if self.thread:
            self.thread.join(timeout=2)
        mhz = self.mhz[getattr(self, "first", 0):]
        out = {"sm_mhz": float(np.median(mhz)) if mhz else None, "sm_max_mhz": self.max_mhz,
               "reasons": sorted(self.reasons), "samples": len(mhz), "source": "NVML, sampled during the timed region"}
        if self.err:
            out["error"] = self.err
        return out


# ------------------------------------------------------------------------------------------
# the B200 arm
# ------------------------------------------------------------------------------------------

def bind_to_gpu_numa_node(index):
    """Run this rank (and so first-touch its pinned staging buffers) on the CPUs NVML reports as
    local to its GPU: with several ranks per box the H2D copies otherwise cross the socket link."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = index
        if vis:
            try:
                idx = int(vis.split(",")[index])
            except ValueError:
                pass
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        return None
    return None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def algorithmic_bytes(batch, unique_reads_per_seg, flags, T, B):
    """B_alg = sum_seg U*C*(8*W_r+4) + sum_c [8*ceil(L_c/32) + 8*(L_c-7)*[pd] + 48] + 8T"""
    W = (batch.read_len + 31) // 32
    ncs = np.diff(batch.seg_contig_start).astype(np.float64)
    pair_bytes = float((unique_reads_per_seg * ncs).sum()) * (8 * W + 4)
    lens = np.diff(batch.contig_off).astype(np.float64)
    per_contig = float((8 * np.ceil(lens / 32)).sum()) + 48.0 * len(lens)
    if flags & B.WANT_PROB_DIST:
        per_contig += float((8 * np.maximum(lens - 7, 0)).sum())
    return pair_bytes, per_contig + 8.0 * T


def unique_reads_per_segment(batch):
    out = np.zeros(batch.n_segments, dtype=np.float64)
    rl = batch.read_len
    for s in range(batch.n_segments):
        r0, r1 = int(batch.seg_read_start[s]), int(batch.seg_read_start[s + 1])
        rows = batch.read_chars[r0 * rl:r1 * rl].reshape(-1, rl)
        out[s] = len(np.unique(np.ascontiguousarray(rows).view(np.dtype((np.void, rl)))))
    return out


def b200_main(args):
    import torch
    import torch.distributed as dist

    from genomeassembler_dev_b200 import breakscore as B

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cpu_base = None
    if world == 1 and not args.no_cpu_baseline:
        # before CUDA is initialised (fork); bounded sample: one segment per host core
        cores = os.cpu_count() or 1
        n_sample = max(cores, 8)
        gbps, rps, t, kind, _ = run_cpu_reference(n_sample, cores)
        cpu_base = {"value": gbps, "unit": UNIT, "cores": cores, "kind": kind, "reads_scored_per_s": rps,
                    "per_core_value": LAST_PER_CORE,
                    "seconds": t,
                    "sample": f"{n_sample} segments of the workload, one process per core, unmodified upstream "
                              f"calc_breakscore (edlib stubbed: off the scored path)"}
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; there is no CPU fallback for the scorer")
    torch.cuda.set_device(local)
    affinity = bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    # ---- inputs: this rank's study ----
    batch = synth.make_batch(args.segments, seed=SEED + rank * args.segments, length=LENGTH, read_len=READ_LEN,
                             coverage=COVERAGE)
    kmers = tables.all_kmer_strings()
    prob = tables.normalised(tables.load_raw())
    sc = B.BreakageScorer(local)
    sc.set_table(kmers, prob)
    # a real (non-default) torch stream: the library launches on it, torch's events and NCCL calls are
    # ordered on it too (handle 0, the legacy default stream, would mean "the context's own stream")
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    sc.set_stream(stream.cuda_stream)
    flags = B.DEFAULT_FLAGS
    Cn, N, S = batch.n_contigs, batch.n_reads, batch.n_segments
    pd_off = B.prob_dist_offsets(batch.contig_off, 8)
    pair_bases = batch.pair_bases()

    # device-resident inputs and outputs
    d_reads = torch.from_numpy(batch.read_chars).to(dev)
    d_ctgs = torch.from_numpy(batch.contig_chars).to(dev)
    d_truth = torch.from_numpy(batch.truth_chars).to(dev)
    d_i32 = torch.zeros(4, Cn, dtype=torch.int32, device=dev)
    d_f64 = torch.zeros(5, Cn, dtype=torch.float64, device=dev)
    d_pd = torch.zeros(max(int(pd_off[-1]), 1), dtype=torch.float64, device=dev)

    def make_batch_struct(rc, cc, tc):
        return B._Batch(S, N, Cn, rc, None, batch.read_len, cc, batch.contig_off.ctypes.data, tc,
                        batch.truth_off.ctypes.data, batch.seg_read_start.ctypes.data, batch.seg_contig_start.ctypes.data)

    def make_result_struct(i32_ptrs, f64_ptrs, pd_ptr):
        r = B._Result()
        r.sequence_len, r.kmer_breaks, r.path_prob_dist_startpos, r.lev_dist_vs_true = i32_ptrs
        (r.bp_score, r.bp_score_norm_by_break_freqs, r.bp_score_norm_by_len, r.ks_stat_prob_dist,
         r.ks_stat_path_freq) = f64_ptrs
        r.path_prob_dist = pd_ptr
        r.path_prob_dist_off = pd_off.ctypes.data
        return r

    db = make_batch_struct(d_reads.data_ptr(), d_ctgs.data_ptr(), d_truth.data_ptr())
    dr = make_result_struct([d_i32[i].data_ptr() for i in range(4)], [d_f64[i].data_ptr() for i in range(5)], d_pd.data_ptr())
    dflags = flags | B.DEVICE_CHARS | B.DEVICE_RESULT

    def barrier():
        if world > 1:
            dist.barrier()

    def gather_records():
        # the path's one exchange: fixed-width per-contig records to rank 0 (NCCL gather over NVLink)
        if world > 1:
            out = [torch.empty_like(d_f64) for _ in range(world)] if rank == 0 else None
            dist.gather(d_f64, out, dst=0)

    def step_device():
        sc.score_batch_raw(db, dr, 8, dflags)
        gather_records()

    sampler = ClockSampler(local)
    sampler.start()  # NVML takes a while to answer the first query: start before the warm-up
    for _ in range(args.warmup):
        step_device()
    torch.cuda.synchronize()
    sc.enable_timing(True)
    sampler.mark()   # keep only samples taken from here on (the timed region)
    launches0 = sc.launch_count
    stage_ms = {k: 0.0 for k in B.STAGES}
    barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    torch.cuda.synchronize()
    barrier()
    for k, v in sc.last_timings().items():  # CUDA events on the launching stream, summed over the K steps
        if v > 0:
            stage_ms[k] += v
    ms_total = e0.elapsed_time(e1)
    launches = sc.launch_count - launches0
    clocks = sampler.stop()
    sc.enable_timing(False)
    t_ms = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    totals = torch.tensor([pair_bases, float(N)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(totals, op=dist.ReduceOp.SUM)
    ms_total = float(t_ms.item())
    all_pair, all_reads = float(totals[0].item()), float(totals[1].item())
    ms_per_step = ms_total / args.steps
    value = all_pair / 1e9 / (ms_per_step / 1e3)

    if args.device_only:  # profiling runs (ncu): the device-resident phase only
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                              "warmup": args.warmup, "ms_per_step": ms_per_step, "device_only": True,
                              "stage_ms_per_step": {k: v / args.steps for k, v in stage_ms.items() if v > 0},
                              "gpu_launches": int(launches), "clocks": clocks}))
        if world > 1:
            dist.destroy_process_group()
        sc.close()
        return 0

    # ---- e2e: host buffers through the C-ABI (pinned), H2D + D2H inside the timed region ----
    def pinned_copy(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t

    h_reads, h_ctgs, h_truth = pinned_copy(batch.read_chars), pinned_copy(batch.contig_chars), pinned_copy(batch.truth_chars)
    h_i32 = torch.zeros(4, Cn, dtype=torch.int32).pin_memory()
    h_f64 = torch.zeros(5, Cn, dtype=torch.float64).pin_memory()
    h_pd = torch.zeros(max(int(pd_off[-1]), 1), dtype=torch.float64).pin_memory()
    hb = make_batch_struct(h_reads.data_ptr(), h_ctgs.data_ptr(), h_truth.data_ptr())
    hr = make_result_struct([h_i32[i].data_ptr() for i in range(4)], [h_f64[i].data_ptr() for i in range(5)], h_pd.data_ptr())
    h2d = int(h_reads.numel() + h_ctgs.numel() + h_truth.numel())
    d2h = int(h_i32[:3].numel() * 4 + h_f64.numel() * 8 + h_pd.numel() * 8)

    def step_host():
        sc.score_batch_raw(hb, hr, 8, flags)  # returns with every result on the host

    for _ in range(max(1, args.warmup // 2)):
        step_host()
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    # the floor of the end-to-end number on this box: the same input bytes through one plain pinned H2D copy
    d_probe = torch.empty_like(d_reads)
    d_probe.copy_(h_reads, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    d_probe.copy_(h_reads, non_blocking=True)
    torch.cuda.synchronize()
    pcie_s = time.perf_counter() - t0
    pcie_gbps = h_reads.numel() / pcie_s / 1e9
    del d_probe
    # one more (untimed) host-buffer step with stage events on, to show where the end-to-end time goes
    sc.enable_timing(True)
    step_host()
    e2e_stage_ms = {k: v for k, v in sc.last_timings().items() if v >= 0}
    sc.enable_timing(False)
    t_e = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_s = float(t_e.item())
    e2e_value = all_pair / 1e9 / (e2e_s / args.steps)
    # device-resident and host paths must agree bit for bit
    same = bool(torch.equal(d_f64.cpu().nan_to_num(nan=-1.0), h_f64.nan_to_num(nan=-1.0)) and
                torch.equal(d_i32[:3].cpu(), h_i32[:3]))

    # ---- the step before the path on the device too: truths and contigs cross PCIe,
    # the reads are simulated on the device (upstream's sampling law) and scored where they are ----
    study = None
    if not args.no_study:
        cap = int(sc._lib.bs_simulate_capacity(batch.truth_off.ctypes.data, S, READ_LEN, COVERAGE))
        d_sim = torch.empty(max(cap, 1), dtype=torch.uint8, device=dev)
        srs = np.zeros(S + 1, np.int64)
        d_tr2 = torch.empty_like(d_truth)
        d_ct2 = torch.empty_like(d_ctgs)

        def step_study():
            d_tr2.copy_(h_truth, non_blocking=True)
            d_ct2.copy_(h_ctgs, non_blocking=True)
            sc._check(sc._lib.bs_simulate_reads(sc._ctx, d_tr2.data_ptr(), batch.truth_off.ctypes.data, S, READ_LEN, COVERAGE, 8,
                                                SEED + rank, B.DEVICE_CHARS | B.DEVICE_RESULT, d_sim.data_ptr(), cap, srs.ctypes.data))
            sb = B._Batch(S, int(srs[-1]), Cn, d_sim.data_ptr(), None, READ_LEN, d_ct2.data_ptr(), batch.contig_off.ctypes.data,
                          d_tr2.data_ptr(), batch.truth_off.ctypes.data, srs.ctypes.data, batch.seg_contig_start.ctypes.data)
            sc.score_batch_raw(sb, hr, 8, flags | B.DEVICE_CHARS)  # results land in the pinned host arrays

        step_study()
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_study()
        torch.cuda.synchronize()
        st_s = time.perf_counter() - t0
        n_s = np.diff(srs).astype(np.float64)
        l_s = np.add.reduceat(np.diff(batch.contig_off), batch.seg_contig_start[:-1]).astype(np.float64) if Cn else np.zeros(S)
        l_s[np.diff(batch.seg_contig_start) == 0] = 0.0
        t_st = torch.tensor([st_s, float((n_s * l_s).sum()), float(n_s.sum())], dtype=torch.float64, device=dev)
        if world > 1:
            tmax = t_st[:1].clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(t_st, op=dist.ReduceOp.SUM)
            t_st[0] = tmax[0]
        study = {"value": float(t_st[1]) / 1e9 / (float(t_st[0]) / args.steps), "unit": UNIT,
                 "ms_per_step": 1e3 * float(t_st[0]) / args.steps, "reads_per_step": float(t_st[2]),
                 "h2d_bytes_per_step": int(h_truth.numel() + h_ctgs.numel()), "d2h_bytes_per_step": d2h,
                 "what": "bs_simulate_reads (upstream's sampling law, device RNG) + bs_score_batch on the device-resident reads; "
                         "only truths and contigs are copied in, every result array is copied out"}

    # ---- roofline: the placement kernel (north_star's hot kernel), every other stage beside it ----
    peak, peak_src = measured_peak()
    uniq = unique_reads_per_segment(batch)
    pair_bytes, other_bytes = algorithmic_bytes(batch, uniq, flags, len(prob), B)
    ms = {k: v / args.steps for k, v in stage_ms.items()}
    W = (batch.read_len + 31) // 32
    lens = np.diff(batch.contig_off).astype(np.float64)
    tlens = np.diff(batch.truth_off).astype(np.float64)
    ascii_in = float(h2d)
    packed_seq = 12.0 * float(np.ceil(lens / 32).sum() + np.ceil(tlens / 32).sum())
    heads = 4.0 * float(sum(max(64, 1 << int(np.ceil(np.log2(max(2 * n, 1))))) for n in np.diff(batch.seg_read_start)))
    w_bytes = 4.0 * float(lens.sum() + len(lens))
    windows = float(np.maximum(lens - 7, 0).sum())
    # bytes each stage has to move at least once (its own formulation), for the per-stage GB/s below
    stage_bytes = {
        "pack": ascii_in + 8.0 * W * N + 4.0 * N + heads + packed_seq,
        "place": 8.0 * W * N + 4.0 * N + heads + 12.0 * float(np.ceil(lens / 32).sum()) + 2 * w_bytes,
        "score": w_bytes + 12.0 * float(np.ceil(lens / 32).sum()) + 48.0 * Cn,
        "truth_spectrum": 12.0 * float(np.ceil(tlens / 32).sum()) + 4.0 * 32896 * S,
        "prob_dist_ks": 8.0 * windows + 12.0 * float(np.ceil(lens / 32).sum()) + 16.0 * Cn,
        "startpos": 12.0 * float(np.ceil(tlens / 32).sum() + np.ceil(lens / 32).sum()) + 8.0 * Cn,
    }
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic_cfg2.json")
    if os.path.exists(tpath):  # dram__bytes_read.sum + dram__bytes_write.sum per launch, one ncu --set full capture
        with open(tpath) as fh:
            tj = json.load(fh)
        if tj.get("segments") == args.segments:
            traffic = tj.get("k_place_index")
    place_ms = ms["place"]
    achieved = pair_bytes / (place_ms / 1e3) / 1e9 if place_ms > 0 else None
    roofline = {
        "kernel": "k_place_index", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": (achieved / peak) if achieved else None, "traffic": traffic, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": pair_bytes, "ms_per_launch": place_ms, "launches_per_step": 1,
        "compulsory": {"bytes_per_launch": stage_bytes["place"],
                       "achieved": stage_bytes["place"] / (place_ms / 1e3) / 1e9 if place_ms > 0 else None,
                       "frac": stage_bytes["place"] / (place_ms / 1e3) / 1e9 / peak if place_ms > 0 else None},
        "stages": {k: {"ms": ms[k], "bytes": stage_bytes[k], "GBps": stage_bytes[k] / (ms[k] / 1e3) / 1e9,
                       "frac_of_peak": stage_bytes[k] / (ms[k] / 1e3) / 1e9 / peak}
                   for k in stage_bytes if ms.get(k, 0) > 0},
        "note": "achieved/frac use SURVEY.md 8(d)'s ALL-PAIRS algorithmic bytes, (8*W_r+4) B per (unique read, contig) "
                "pair; the index formulation never touches most pairs, so frac > 1 is not a DRAM rate. "
                "'compulsory' = bytes this launch must move at least once / its time: the kernel's real HBM "
                "efficiency (it is L2-latency and issue bound); 'traffic' = ncu DRAM bytes per launch.",
    }

    # ---- both table passes of an experiment (real, then uniform) from ONE placement ----
    two_pass = None
    if rank == 0 and not args.no_study:
        sc.set_second_table(tables.uniform(len(prob)))
        d_f64b = torch.zeros(5, Cn, dtype=torch.float64, device=dev)
        d_pd2 = torch.zeros_like(d_pd)
        dr2 = make_result_struct([d_i32[i].data_ptr() for i in range(4)], [d_f64[i].data_ptr() for i in range(5)], d_pd.data_ptr())
        (dr2.bp_score2, dr2.bp_score_norm_by_break_freqs2, dr2.bp_score_norm_by_len2, dr2.ks_stat_prob_dist2,
         dr2.ks_stat_path_freq2) = [d_f64b[i].data_ptr() for i in range(5)]
        dr2.path_prob_dist2 = d_pd2.data_ptr()
        for _ in range(2):
            sc.score_batch_raw(db, dr2, 8, dflags | B.WANT_SECOND_TABLE)
        torch.cuda.synchronize()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(args.steps):
            sc.score_batch_raw(db, dr2, 8, dflags | B.WANT_SECOND_TABLE)
        f1.record()
        torch.cuda.synchronize()
        both_ms = f0.elapsed_time(f1) / args.steps
        two_pass = {"ms_per_step_both_tables_one_call": both_ms, "ms_per_step_one_table": ms_per_step,
                    "vs_two_separate_calls": both_ms / (2 * ms_per_step),
                    "what": "real table + uniform table (the R driver's two passes, lib/DeNovoAssembler.R:325-333) scored "
                            "from one packing/placement/startpos; device-resident, rank 0"}
        sc.set_second_table(None)

    # ---- for the record: north_star's all-pairs kernel (BS_PLACE_SCAN) on a sample of the workload ----
    scan = None
    if rank == 0 and args.scan_segments > 0:
        ns = min(args.scan_segments, S)
        r1, c1 = int(batch.seg_read_start[ns]), int(batch.seg_contig_start[ns])
        sb = B._Batch(ns, r1, c1, d_reads.data_ptr(), None, batch.read_len, d_ctgs.data_ptr(), batch.contig_off.ctypes.data,
                      d_truth.data_ptr(), batch.truth_off.ctypes.data, batch.seg_read_start.ctypes.data,
                      batch.seg_contig_start.ctypes.data)
        scan_ms = {}
        for name, fl in (("all_pairs_scan", B.PLACE_SCAN), ("contig_tile_index", B.PLACE_TILE), ("read_index", 0)):
            sc.score_batch_raw(sb, dr, 8, B.DEVICE_CHARS | B.DEVICE_RESULT | fl)
            sc.enable_timing(True)
            sc.score_batch_raw(sb, dr, 8, B.DEVICE_CHARS | B.DEVICE_RESULT | fl)
            scan_ms[name] = sc.last_timings()["place"]
            sc.enable_timing(False)
        sample_bytes = float((uniq[:ns] * np.diff(batch.seg_contig_start)[:ns]).sum()) * (8 * W + 4)
        scan = {"sample_segments": ns, "algorithmic_bytes": sample_bytes,
                "place_ms": scan_ms,
                "achieved_GBps": {k: sample_bytes / (v / 1e3) / 1e9 for k, v in scan_ms.items() if v > 0},
                "frac_of_peak": {k: sample_bytes / (v / 1e3) / 1e9 / peak for k, v in scan_ms.items() if v > 0},
                "note": "same SURVEY 8(d) algorithmic bytes for the three placement kernels on the first segments of the "
                        "workload: the all-pairs scan is what north_star describes (every read against every contig position)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8 compare / int32 count / f64 sums", "data": "synthetic",
            "reads_scored_per_s": all_reads / (ms_per_step / 1e3),
            "config": {"workload": workload_name(args.segments), "segments_per_gpu": args.segments,
                       "reads_per_gpu": N, "contigs_per_gpu": Cn, "outputs": "scores+kmer_breaks+startpos+path_prob_dist+KS",
                       "l2": "inputs per step (%.0f MB ASCII) exceed the 126 MB L2; no flush" % (h2d / 1e6),
                       "parallelism": f"segments sharded over {world} GPU(s), NCCL gather of score records",
                       "cpu_affinity": affinity},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * e2e_s / args.steps, "matches_device_resident_run": same,
                    "stage_ms_untimed_extra_step": e2e_stage_ms,
                    "plain_pinned_h2d_GBps": pcie_gbps, "h2d_floor_ms": 1e3 * h2d / (pcie_gbps * 1e9)},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
        }
        if study:
            line["study_with_device_simulated_reads"] = study
        if scan:
            line["placement_variants"] = scan
        if two_pass:
            line["two_table_passes"] = two_pass
        if cpu_base:
            line["cpu_baseline"] = cpu_base
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    sc.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--segments", type=int, default=1000, help="segments per GPU (cfg-2 study size: 1000)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--device-only", action="store_true", help="device-resident phase only (profiling runs)")
    ap.add_argument("--no-study", action="store_true", help="skip the extra simulate-on-device measurement")
    ap.add_argument("--scan-segments", type=int, default=20, help="segments timed with the all-pairs / tile placement kernels (0: skip)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        print("bench.py: warning: fewer than 3 warm-up steps", file=sys.stderr)
    if args.impl == "reference":
        return reference_main(args)
    return b200_main(args)


if __name__ == "__main__":
    sys.exit(main())
