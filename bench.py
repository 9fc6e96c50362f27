#!/usr/bin/env python
"""Benchmark of the breakage-scoring hot path (BASELINE.json metric: reads scored/s and
read x contig Gbp compared/s; % of HBM peak).

Headline line (the contract's keys): a "step" is one pass of the hot path over one batch -- the cfg-2 study of
BASELINE.json (`--segments` synthetic 50 kb segments, 150 bp reads at 30x, velvet-style contig sets) scored by ONE
bs_score_batch call per rank with the upstream default outputs (scores, kmer_breaks, startpos, path_prob_dist, KS
statistics).  Ranks own disjoint studies (weak scaling: segments are independent units, no data-path collective;
per-contig score records are gathered to rank 0 over NCCL).

  value     whole-job Gbp/s, inputs already resident in HBM (device pointers through the C-ABI)
  e2e       the same metric through the C-ABI with HOST buffers: H2D of the ASCII inputs and D2H of
            every result array inside the timed region
  roofline  the slowest stage's kernel: max(bytes it must move at least once, ncu DRAM bytes) / its CUDA-event
            duration, against MEASURED_PEAKS.json hbm_gbs; the whole step the same way; SURVEY.md 8(d)'s ALL-PAIRS
            equivalent is a separate key (the index formulation never touches most pairs, so it is not a DRAM rate)
  cpu_baseline / --impl reference: the UNMODIFIED upstream calc_breakscore (oracle/_ref, edit distance
            stubbed: off the scored path) on the box's host cores, one process per core over
            disjoint segments of the same workload (bounded sample, rate-normalised)

Extra keys of the same line (measured in the same run, after the headline):
  strong_scaling   the SAME 1000-segment study split over the N ranks (N > 1)
  contig_sharded   BASELINE.json configs[3] and [4] with north_star's multi-GPU split: ONE segment, its contigs dealt
                   out over the ranks longest-first, reads replicated, fixed-width records gathered over NCCL --
                   cfg4 (10 000 scaffolds) and cfg5 (100 Mb truth, 1e8 reads, 1e5 contigs, generated on the device);
                   each with device-resident and host-to-host timings, the cost of replicating the reads two ways,
                   a sampled oracle diff per rank and the gathered table compared with one GPU scoring everything
                   (tools/contig_sharded_bench.py).  cfg4 also carries `compositional`: the same scaffolds given as parts
                   of their 16 base contigs and scored from the parts (bs_score_scaffolds), sharded the same way, with
                   its own oracle sample and every column compared with the rescan of the texts.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(1, os.path.join(ROOT, "tools"))

from genomeassembler_dev_b200 import sharding, synth, tables  # noqa: E402

LAST_PER_CORE = None
METRIC = "read_x_contig_Gbp_compared_per_s"
UNIT = "Gbp/s"
DTYPE = "u8 compare / int32 count / f64 sums"
SEED = 1234
LENGTH, READ_LEN, COVERAGE = 50_000, 150, 30.0
REF_SEGMENTS_PER_CORE = 4

# kernels of every pipeline stage (per-kernel ncu DRAM bytes are summed per stage: profiles/traffic_cfg2.json)
STAGE_KERNELS = {
    "pack": ["k_pack_seqs", "k_pack_reads_bulk", "k_pack_reads_uniform"],
    "place": ["k_place_index"],
    "score": ["k_break_score"],
    "truth_spectrum": ["k_truth_spectrum_smem"],
    "prob_dist_ks": ["k_prob_dist_ks_small", "k_prob_dist_ks"],
    "startpos": ["k_startpos_build", "k_startpos_scan", "k_startpos_verify", "k_startpos"],
}
STAGE_MAIN_KERNEL = {"pack": "k_pack_reads_bulk", "place": "k_place_index", "score": "k_break_score",
                     "truth_spectrum": "k_truth_spectrum_smem", "prob_dist_ks": "k_prob_dist_ks", "startpos": "k_startpos_scan"}
STAGE_LIMITER = {
    "pack": "streaming: DRAM + issue",
    "place": "divergent gathers (L1 wavefronts) + dependent L2 round trips: bucket head -> chain entry -> packed read",
    "score": "divergent table gathers + shared-memory atomics",
    "truth_spectrum": "shared-memory atomics, output writes",
    "prob_dist_ks": "divergent gathers (window table entry, truth counts); with kmer == 8 also the break scoring (position weights, KS-B hash)",
    "startpos": "shared-memory bitmap probes per truth position (issue)",
}


def workload_name(n_segments):
    return (f"cfg2: {n_segments} synthetic 50 kb segments per GPU, 150 bp reads 30x, velvet-style contig sets "
            f"(5-60 contigs/segment, 10% mutated), kmer=8, real breakage table")


def study_config(n_segments, n_gpus):
    """the SAME dict in both arms (b200 / reference): nothing measured goes in here"""
    return {
        "workload": workload_name(n_segments), "segments_per_gpu": n_segments,
        "outputs": "scores+kmer_breaks+startpos+path_prob_dist+KS",
        "l2": "per-step inputs (about 1.6 GB of ASCII per 1000 segments) exceed the 126 MB L2; no flush",
        "parallelism": f"segments sharded over {n_gpus} GPU(s), NCCL gather of score records",
        "reference_arm": f"rate-normalised: every step of --impl reference scores a bounded sample of this workload "
                         f"({REF_SEGMENTS_PER_CORE} segments per host core, one process per core) and reports Gbp/s of that sample",
    }


# ------------------------------------------------------------------------------------------
# CPU checkers (test/bench infrastructure: oracle/_ref = unmodified upstream code, oracle/ = C restatement).
# Everything below runs in forked worker processes BEFORE CUDA is initialised in this one.
# ------------------------------------------------------------------------------------------

_POOL = {}  # data the forked workers read (set in the parent before the pool is created)


def _table():
    if "kmers" not in _POOL:
        _POOL["kmers"] = tables.all_kmer_strings()
        _POOL["prob"] = tables.normalised(tables.load_raw())
    return _POOL["kmers"], _POOL["prob"]


def _ref_task(args):
    seeds, use_ref = args
    from oracle import loader as O
    kmers, prob = _table()
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


def load_checkers():
    """dlopen the CPU checkers in THIS process before forking: the workers inherit the mappings, and the driver's
    record of the libraries this arm loaded shows the unmodified reference build"""
    from oracle import loader as O
    O.build()
    _table()
    use_ref = O.have_ref()
    if use_ref:
        O._load(os.path.join("_ref", "libref_breakscore_noedit.so"))
    O._load("liboracle.so")
    return O, use_ref


def run_cpu_reference(n_segments_sample, cores, seed0=SEED):
    """one process per core, disjoint segments; returns (Gbp/s, reads/s, seconds, kind, wall)"""
    import multiprocessing as mp
    _, use_ref = load_checkers()
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


def _contig_task(args):
    """checker on a few contigs of one segment against ALL of its reads: 'oracle' -> every record column incl. the
    KS statistics (C restatement); 'time' -> the unmodified reference where built, timed"""
    key, idx, mode = args
    from oracle import loader as O
    kmers, prob = _table()
    d = _POOL[key]
    ctgs = [d["contigs"][i] for i in idx]
    t0 = time.perf_counter()
    if mode == "time" and O.have_ref():
        O.ref_calc_breakscore(ctgs, d["read_list"], d["truth"], 8, kmers, prob, edit_distance=False, want_prob_dist=False)
        return time.perf_counter() - t0, float(len(d["read_list"])) * float(sum(len(c) for c in ctgs)), "reference"
    o = O.oracle_calc_breakscore(ctgs, d["read_list"], d["truth"], 8, kmers, prob, want_prob_dist=False)
    if mode == "time":
        return time.perf_counter() - t0, float(len(d["read_list"])) * float(sum(len(c) for c in ctgs)), "port"
    return {k: np.asarray(o[k]) for k in sharding.RECORD_F64 + sharding.RECORD_I32}


def pool_map(fn, tasks, procs):
    import multiprocessing as mp
    if not tasks:
        return []
    with mp.get_context("fork").Pool(max(1, min(procs, len(tasks)))) as pool:
        return pool.map(fn, tasks)


def oracle_sample(key, sample_idx, procs):
    """expected record columns of the contigs sample_idx of _POOL[key] (C restatement of the reference, all reads)"""
    groups = [sample_idx[i:i + 2] for i in range(0, len(sample_idx), 2)]
    res = pool_map(_contig_task, [(key, [int(x) for x in g], "oracle") for g in groups], procs)
    if not res:
        return {}
    return {k: np.concatenate([r[k] for r in res]) for k in res[0]}


def cpu_rate(key, idx_groups, procs):
    """Gbp/s of the CPU reference on a sample of _POOL[key]'s contigs (all reads), one process per group"""
    res = pool_map(_contig_task, [(key, [int(x) for x in g], "time") for g in idx_groups], procs)
    t = max(r[0] for r in res)
    return {"value": sum(r[1] for r in res) / 1e9 / t, "unit": UNIT, "cores": min(procs, len(idx_groups)), "kind": res[0][2],
            "seconds": t, "per_core_value": float(np.mean([r[1] / 1e9 / r[0] for r in res]))}


def pick_sample(parts_of_rank, contigs, n, seed):
    """n contigs of a rank's share: the longest, both members of a duplicate pair if there is one, the rest at random"""
    mine = np.asarray(parts_of_rank, dtype=np.int64)
    if len(mine) <= n:
        return mine
    rng = np.random.default_rng(seed)
    lens = np.array([len(contigs[i]) for i in mine])
    chosen = {int(mine[int(np.argmax(lens))])}
    seen = {}
    for i in mine:
        c = contigs[int(i)]
        if c in seen:
            chosen.update((int(i), seen[c]))
            break
        seen[c] = int(i)
    rest = [int(i) for i in rng.permutation(mine) if int(i) not in chosen]
    chosen.update(rest[:max(0, n - len(chosen))])
    return np.array(sorted(chosen), dtype=np.int64)


def reference_main(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    per_step = REF_SEGMENTS_PER_CORE * cores
    for _ in range(args.warmup):
        run_cpu_reference(cores, cores)
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
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": DTYPE,
        "data": "synthetic", "reads_scored_per_s": tot_reads / tot_t,
        "config": study_config(args.segments, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "per_core_value": LAST_PER_CORE,
                         "sample": f"{per_step} segments of the workload per step ({REF_SEGMENTS_PER_CORE} per core), one process per "
                                   f"core, unmodified upstream calc_breakscore (edlib stubbed); rate = bases compared / time of "
                                   f"the slowest process"},
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


def all_pairs_bytes(batch, unique_reads_per_seg):
    """sum_seg U*C*(8*W_r+4) -- what an all-pairs formulation would stream"""
    W = (batch.read_len + 31) // 32
    ncs = np.diff(batch.seg_contig_start).astype(np.float64)
    return float((unique_reads_per_seg * ncs).sum()) * (8 * W + 4)


def unique_reads_per_segment(batch):
    out = np.zeros(batch.n_segments, dtype=np.float64)
    rl = batch.read_len
    for s in range(batch.n_segments):
        r0, r1 = int(batch.seg_read_start[s]), int(batch.seg_read_start[s + 1])
        rows = batch.read_chars[r0 * rl:r1 * rl].reshape(-1, rl)
        out[s] = len(np.unique(np.ascontiguousarray(rows).view(np.dtype((np.void, rl)))))
    return out


def compulsory_stage_bytes(batch, h2d_bytes):
    """bytes every stage has to move at least once in ITS OWN formulation"""
    N, Cn, S = batch.n_reads, batch.n_contigs, batch.n_segments
    W = (batch.read_len + 31) // 32
    lens = np.diff(batch.contig_off).astype(np.float64)
    tlens = np.diff(batch.truth_off).astype(np.float64)
    cw, tw = float(np.ceil(lens / 32).sum()), float(np.ceil(tlens / 32).sum())
    heads = 4.0 * float(sum(max(64, 1 << int(np.ceil(np.log2(max(2 * n, 1))))) for n in np.diff(batch.seg_read_start)))
    w_bytes = 4.0 * float(lens.sum() + len(lens))
    windows = float(np.maximum(lens - 7, 0).sum())
    return {
        "pack": float(h2d_bytes) + 8.0 * W * N + 8.0 * N + N + heads + 12.0 * (cw + tw),   # ASCII in; words, chain entries, flags, bucket heads out
        "place": 8.0 * W * N + 8.0 * N + heads + 12.0 * cw + 2 * w_bytes,
        "score": w_bytes + 12.0 * cw + 48.0 * Cn,
        "truth_spectrum": 12.0 * tw + (4.0 + 8.0) * 32897 * S,                            # cumulative counts + per-x-value counts out
        "prob_dist_ks": 8.0 * windows + 12.0 * cw + 16.0 * Cn,
        "startpos": 12.0 * (tw + cw) + 8.0 * Cn,
    }


def load_traffic(n_segments):
    """ncu dram__bytes_read.sum + dram__bytes_write.sum per launch and kernel (profiles/traffic_cfg2.json, regenerated
    from the round's --set full capture by tools/traffic_from_ncu.py), summed per stage; None when absent"""
    tpath = os.path.join(ROOT, "profiles", "traffic_cfg2.json")
    if not os.path.exists(tpath):
        return None, None
    with open(tpath) as fh:
        tj = json.load(fh)
    if tj.get("segments") != n_segments:
        return None, None
    per_stage = {}
    for st, kernels in STAGE_KERNELS.items():
        vals = [float(tj[k]) for k in kernels if k in tj]
        per_stage[st] = sum(vals) if vals else None
    return per_stage, tj.get("source")


def build_roofline(batch, ms, h2d_bytes, n_segments, all_pairs):
    peak, peak_src = measured_peak()
    comp = compulsory_stage_bytes(batch, h2d_bytes)
    traffic, traffic_src = load_traffic(n_segments)
    if ms.get("score", 0.0) <= 0:  # kmer == 8: the KS-A kernels score on their way over the windows (no k_break_score launch)
        comp["prob_dist_ks"] += comp.pop("score")
        if traffic and traffic.get("score") and traffic.get("prob_dist_ks"):
            traffic["prob_dist_ks"] += traffic.pop("score")
    stages = {}
    for k, b in comp.items():
        t = ms.get(k, 0.0)
        if t <= 0:
            continue
        tr = traffic.get(k) if traffic else None
        phys = max(b, tr) if tr else b
        stages[k] = {"ms": t, "compulsory_bytes": b, "traffic": tr, "GBps": phys / (t / 1e3) / 1e9,
                     "frac": phys / (t / 1e3) / 1e9 / peak, "frac_compulsory": b / (t / 1e3) / 1e9 / peak,
                     "kernel": STAGE_MAIN_KERNEL[k], "limiter": STAGE_LIMITER[k]}
    slow = max(stages, key=lambda k: stages[k]["ms"])
    step_ms = sum(v["ms"] for v in stages.values())
    step_comp = sum(v["compulsory_bytes"] for v in stages.values())
    step_tr = sum(v["traffic"] for v in stages.values()) if traffic and all(v["traffic"] for v in stages.values()) else None
    s = stages[slow]
    return {
        "kernel": s["kernel"], "stage": slow, "bound": "hbm", "limiter": s["limiter"],
        "achieved": s["GBps"], "peak": peak, "unit": "GB/s", "frac": s["frac"], "traffic": s["traffic"],
        "compulsory_bytes_per_launch": s["compulsory_bytes"], "ms_per_launch": s["ms"], "launches_per_step": 1,
        "peak_source": peak_src, "traffic_source": traffic_src,
        "definition": "the step's slowest stage; achieved = max(bytes the launch must move at least once, ncu DRAM bytes) / its "
                      "CUDA-event time on the launching stream; frac = achieved / peak",
        "step": {"ms_kernels": step_ms, "compulsory_bytes": step_comp, "traffic": step_tr,
                 "frac_compulsory": step_comp / (step_ms / 1e3) / 1e9 / peak,
                 "frac_traffic": (step_tr / (step_ms / 1e3) / 1e9 / peak) if step_tr else None,
                 "floor_ms_at_peak": step_comp / (peak * 1e9) * 1e3},
        "stages": stages,
        "all_pairs_equivalent": {"bytes_per_step": all_pairs,
                                 "GBps_over_place_time": all_pairs / (ms["place"] / 1e3) / 1e9 if ms.get("place", 0) > 0 else None,
                                 "note": "SURVEY.md 8(d): (8*W_r+4) B per (unique read, contig) pair of an ALL-PAIRS formulation. The read-index "
                                         "placement never touches most pairs, so this is work avoided, not a DRAM rate: no frac."},
    }


class Rig:
    """what every measurement below needs: the scorer, torch, the process group, this rank's stream"""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from genomeassembler_dev_b200 import breakscore as B
        self.torch, self.dist, self.B, self.args = torch, dist, B, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; there is no CPU fallback for the scorer")
        torch.cuda.set_device(self.local)
        self.affinity = bind_to_gpu_numa_node(self.local) if self.world > 1 else None
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.kmers, self.prob = _table()
        self.sc = B.BreakageScorer(self.local)
        self.sc.set_table(self.kmers, self.prob)
        # a real (non-default) torch stream: the library launches on it, torch's events and NCCL calls are
        # ordered on it too (handle 0, the legacy default stream, would mean "the context's own stream")
        self.stream = torch.cuda.Stream(device=self.dev)
        torch.cuda.set_stream(self.stream)
        assert self.stream.cuda_stream != 0
        self.sc.set_stream(self.stream.cuda_stream)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def sync(self):
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, xs):
        t = self.torch.tensor(list(xs), dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [float(v) for v in t.tolist()]

    def pinned(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(a)).pin_memory()

    def time_device(self, fn, steps, warmup):
        """ms per step: CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks"""
        torch = self.torch
        for _ in range(warmup):
            fn()
        self.sync()
        self.barrier()
        self.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.sync()
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1)) / steps

    def time_wall(self, fn, steps, warmup):
        """ms per step by wall clock around synchronised steps (host-to-host paths), max over ranks"""
        for _ in range(warmup):
            fn()
        self.sync()
        self.barrier()
        self.sync()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        self.sync()
        dt = time.perf_counter() - t0
        self.barrier()
        return self.max_over_ranks(1e3 * dt) / steps


class StudyBuffers:
    """device-resident and pinned-host argument structs of one batch of segments"""

    def __init__(self, rig, batch):
        torch, self.B = rig.torch, rig.B
        self.batch = batch
        Cn = batch.n_contigs
        self.pd_off = self.B.prob_dist_offsets(batch.contig_off, 8)
        dev = rig.dev
        self.d_in = [torch.from_numpy(x).to(dev) for x in (batch.read_chars, batch.contig_chars, batch.truth_chars)]
        self.d_i32 = torch.zeros(4, max(Cn, 1), dtype=torch.int32, device=dev)
        self.d_f64 = torch.zeros(5, max(Cn, 1), dtype=torch.float64, device=dev)
        self.d_pd = torch.zeros(max(int(self.pd_off[-1]), 1), dtype=torch.float64, device=dev)
        self.db = self.batch_struct(*[t.data_ptr() for t in self.d_in])
        self.dr = self.result_struct(self.d_i32, self.d_f64, self.d_pd)
        self.h_in = None

    def batch_struct(self, rc, cc, tc):
        b = self.batch
        return self.B._Batch(b.n_segments, b.n_reads, b.n_contigs, rc, None, b.read_len, cc, b.contig_off.ctypes.data, tc,
                             b.truth_off.ctypes.data, b.seg_read_start.ctypes.data, b.seg_contig_start.ctypes.data)

    def result_struct(self, i32, f64, pd):
        r = self.B._Result()
        r.sequence_len, r.kmer_breaks, r.path_prob_dist_startpos, r.lev_dist_vs_true = [i32[i].data_ptr() for i in range(4)]
        (r.bp_score, r.bp_score_norm_by_break_freqs, r.bp_score_norm_by_len, r.ks_stat_prob_dist,
         r.ks_stat_path_freq) = [f64[i].data_ptr() for i in range(5)]
        r.path_prob_dist = pd.data_ptr()
        r.path_prob_dist_off = self.pd_off.ctypes.data
        return r

    def make_host(self, rig):
        torch, b = rig.torch, self.batch
        Cn = b.n_contigs
        self.h_in = [rig.pinned(x) for x in (b.read_chars, b.contig_chars, b.truth_chars)]
        self.h_i32 = torch.zeros(4, max(Cn, 1), dtype=torch.int32).pin_memory()
        self.h_f64 = torch.zeros(5, max(Cn, 1), dtype=torch.float64).pin_memory()
        self.h_pd = torch.zeros(max(int(self.pd_off[-1]), 1), dtype=torch.float64).pin_memory()
        self.hb = self.batch_struct(*[t.data_ptr() for t in self.h_in])
        self.hr = self.result_struct(self.h_i32, self.h_f64, self.h_pd)
        self.h2d = int(sum(t.numel() for t in self.h_in))
        self.d2h = int(self.h_i32[:3].numel() * 4 + self.h_f64.numel() * 8 + self.h_pd.numel() * 8)


def b200_main(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    cores = os.cpu_count() or 1
    procs = max(1, cores // world)
    me = sys.modules[__name__]

    # ---- CPU work first (forked pools; CUDA is not initialised yet) ----
    cpu_base = None
    if world == 1 and not args.no_cpu_baseline and not args.device_only:
        n_sample = REF_SEGMENTS_PER_CORE * cores
        gbps, rps, t, kind, _ = run_cpu_reference(n_sample, cores)
        cpu_base = {"value": gbps, "unit": UNIT, "cores": cores, "kind": kind, "reads_scored_per_s": rps,
                    "per_core_value": LAST_PER_CORE, "seconds": t,
                    "sample": f"{n_sample} segments of the workload ({REF_SEGMENTS_PER_CORE} per core), one process per core, unmodified "
                              f"upstream calc_breakscore (edlib stubbed: off the scored path)"}
    sharded_plan = None
    if not args.skip_sharded and not args.device_only:
        import contig_sharded_bench as CS  # tools/: the cfg-4 / cfg-5 measurements
        sharded_plan = CS.prepare_cpu_side(me, args, rank, world, procs)

    rig = Rig(args)
    torch, dist, B, sc, dev = rig.torch, rig.dist, rig.B, rig.sc, rig.dev

    # ---- inputs: this rank's study ----
    batch = synth.make_batch(args.segments, seed=SEED + rank * args.segments, length=LENGTH, read_len=READ_LEN,
                             coverage=COVERAGE)
    flags = B.DEFAULT_FLAGS
    Cn, N = batch.n_contigs, batch.n_reads
    pair_bases = batch.pair_bases()
    sb = StudyBuffers(rig, batch)
    dflags = flags | B.DEVICE_CHARS | B.DEVICE_RESULT

    def gather_records():
        # the path's one exchange: fixed-width per-contig records to rank 0 (NCCL gather over NVLink)
        if world > 1:
            out = [torch.empty_like(sb.d_f64) for _ in range(world)] if rank == 0 else None
            dist.gather(sb.d_f64, out, dst=0)

    def step_device():
        sc.score_batch_raw(sb.db, sb.dr, 8, dflags)
        gather_records()

    sampler = ClockSampler(rig.local)
    sampler.start()  # NVML takes a while to answer the first query: start before the warm-up
    for _ in range(args.warmup):
        step_device()
    rig.sync()
    sc.enable_timing(True)
    sampler.mark()   # keep only samples taken from here on (the timed region)
    launches0 = sc.launch_count
    rig.barrier()
    rig.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    rig.sync()
    rig.barrier()
    stage_ms = {k: v for k, v in sc.last_timings().items() if v > 0}  # CUDA events on the launching stream, summed over the K steps
    launches = sc.launch_count - launches0
    clocks = sampler.stop()
    sc.enable_timing(False)
    ms_per_step = rig.max_over_ranks(e0.elapsed_time(e1)) / args.steps
    all_pair, all_reads = rig.sum_over_ranks([pair_bases, float(N)])
    value = all_pair / 1e9 / (ms_per_step / 1e3)
    ms = {k: v / args.steps for k, v in stage_ms.items()}

    if args.device_only:  # profiling runs (ncu): the device-resident phase only
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                              "warmup": args.warmup, "ms_per_step": ms_per_step, "device_only": True,
                              "stage_ms_per_step": ms, "gpu_launches": int(launches), "clocks": clocks}))
        if world > 1:
            dist.destroy_process_group()
        sc.close()
        return 0

    # ---- e2e: host buffers through the C-ABI (pinned), H2D + D2H inside the timed region ----
    sb.make_host(rig)
    h2d, d2h = sb.h2d, sb.d2h

    def step_host():
        sc.score_batch_raw(sb.hb, sb.hr, 8, flags)  # returns with every result on the host

    e2e_ms = rig.time_wall(step_host, args.steps, max(1, args.warmup // 2))
    # the floor of the end-to-end number on this box: the same input bytes through one plain pinned H2D copy
    d_probe = torch.empty_like(sb.d_in[0])
    d_probe.copy_(sb.h_in[0], non_blocking=True)
    rig.sync()
    t0 = time.perf_counter()
    d_probe.copy_(sb.h_in[0], non_blocking=True)
    rig.sync()
    pcie_gbps = sb.h_in[0].numel() / (time.perf_counter() - t0) / 1e9
    del d_probe
    # one more (untimed) host-buffer step with stage events on, to show where the end-to-end time goes
    sc.enable_timing(True)
    step_host()
    e2e_stage_ms = {k: v for k, v in sc.last_timings().items() if v >= 0}
    sc.enable_timing(False)
    e2e_value = all_pair / 1e9 / (e2e_ms / 1e3)
    # device-resident and host paths must agree bit for bit
    same = bool(torch.equal(sb.d_f64.cpu().nan_to_num(nan=-1.0), sb.h_f64.nan_to_num(nan=-1.0)) and
                torch.equal(sb.d_i32[:3].cpu(), sb.h_i32[:3]))

    # ---- the step before the path on the device too: truths and contigs cross PCIe,
    # the reads are simulated on the device (upstream's sampling law) and scored where they are ----
    study = None
    S = batch.n_segments
    if not args.no_study:
        cap = int(sc._lib.bs_simulate_capacity(batch.truth_off.ctypes.data, S, READ_LEN, COVERAGE))
        d_sim = torch.empty(max(cap, 1), dtype=torch.uint8, device=dev)
        srs = np.zeros(S + 1, np.int64)
        d_tr2, d_ct2 = torch.empty_like(sb.d_in[2]), torch.empty_like(sb.d_in[1])

        def step_study():
            d_tr2.copy_(sb.h_in[2], non_blocking=True)
            d_ct2.copy_(sb.h_in[1], non_blocking=True)
            sc._check(sc._lib.bs_simulate_reads(sc._ctx, d_tr2.data_ptr(), batch.truth_off.ctypes.data, S, READ_LEN, COVERAGE, 8,
                                                SEED + rank, B.DEVICE_CHARS | B.DEVICE_RESULT, d_sim.data_ptr(), cap, srs.ctypes.data))
            b2 = B._Batch(S, int(srs[-1]), Cn, d_sim.data_ptr(), None, READ_LEN, d_ct2.data_ptr(), batch.contig_off.ctypes.data,
                          d_tr2.data_ptr(), batch.truth_off.ctypes.data, srs.ctypes.data, batch.seg_contig_start.ctypes.data)
            sc.score_batch_raw(b2, sb.hr, 8, flags | B.DEVICE_CHARS)  # results land in the pinned host arrays

        st_ms = rig.time_wall(step_study, args.steps, 1)
        n_s = np.diff(srs).astype(np.float64)
        l_s = np.add.reduceat(np.diff(batch.contig_off), batch.seg_contig_start[:-1]).astype(np.float64) if Cn else np.zeros(S)
        l_s[np.diff(batch.seg_contig_start) == 0] = 0.0
        tot = rig.sum_over_ranks([float((n_s * l_s).sum()), float(n_s.sum())])
        study = {"value": tot[0] / 1e9 / (st_ms / 1e3), "unit": UNIT, "ms_per_step": st_ms, "reads_per_step": tot[1],
                 "h2d_bytes_per_step": int(sb.h_in[2].numel() + sb.h_in[1].numel()), "d2h_bytes_per_step": d2h,
                 "what": "bs_simulate_reads (upstream's sampling law, device RNG) + bs_score_batch on the device-resident reads; "
                         "only truths and contigs are copied in, every result array is copied out"}
        del d_sim, d_tr2, d_ct2

    # ---- roofline ----
    uniq = unique_reads_per_segment(batch)
    roofline = build_roofline(batch, ms, h2d, args.segments, all_pairs_bytes(batch, uniq))

    # ---- both table passes of an experiment (real, then uniform) from ONE placement ----
    two_pass = None
    if rank == 0 and not args.no_study:
        sc.set_second_table(tables.uniform(len(rig.prob)))
        d_f64b = torch.zeros(5, Cn, dtype=torch.float64, device=dev)
        d_pd2 = torch.zeros_like(sb.d_pd)
        dr2 = sb.result_struct(sb.d_i32, sb.d_f64, sb.d_pd)
        (dr2.bp_score2, dr2.bp_score_norm_by_break_freqs2, dr2.bp_score_norm_by_len2, dr2.ks_stat_prob_dist2,
         dr2.ks_stat_path_freq2) = [d_f64b[i].data_ptr() for i in range(5)]
        dr2.path_prob_dist2 = d_pd2.data_ptr()
        for _ in range(2):
            sc.score_batch_raw(sb.db, dr2, 8, dflags | B.WANT_SECOND_TABLE)
        rig.sync()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(args.steps):
            sc.score_batch_raw(sb.db, dr2, 8, dflags | B.WANT_SECOND_TABLE)
        f1.record()
        rig.sync()
        both_ms = f0.elapsed_time(f1) / args.steps
        two_pass = {"ms_per_step_both_tables_one_call": both_ms, "ms_per_step_one_table": ms_per_step,
                    "vs_two_separate_calls": both_ms / (2 * ms_per_step),
                    "what": "real table + uniform table (the R driver's two passes, lib/DeNovoAssembler.R:325-333) scored "
                            "from one packing/placement/startpos; device-resident, rank 0"}
        sc.set_second_table(None)
        del d_f64b, d_pd2

    # ---- for the record: north_star's all-pairs kernel (BS_PLACE_SCAN) on a sample of the workload ----
    scan = None
    if rank == 0 and args.scan_segments > 0:
        W = (batch.read_len + 31) // 32
        ns = min(args.scan_segments, S)
        r1, c1 = int(batch.seg_read_start[ns]), int(batch.seg_contig_start[ns])
        b3 = B._Batch(ns, r1, c1, sb.d_in[0].data_ptr(), None, batch.read_len, sb.d_in[1].data_ptr(), batch.contig_off.ctypes.data,
                      sb.d_in[2].data_ptr(), batch.truth_off.ctypes.data, batch.seg_read_start.ctypes.data,
                      batch.seg_contig_start.ctypes.data)
        scan_ms = {}
        for name, fl in (("all_pairs_scan", B.PLACE_SCAN), ("contig_tile_index", B.PLACE_TILE), ("read_index", 0)):
            sc.score_batch_raw(b3, sb.dr, 8, B.DEVICE_CHARS | B.DEVICE_RESULT | fl)
            sc.enable_timing(True)
            sc.score_batch_raw(b3, sb.dr, 8, B.DEVICE_CHARS | B.DEVICE_RESULT | fl)
            scan_ms[name] = sc.last_timings()["place"]
            sc.enable_timing(False)
        sample_bytes = float((uniq[:ns] * np.diff(batch.seg_contig_start)[:ns]).sum()) * (8 * W + 4)
        scan = {"sample_segments": ns, "all_pairs_bytes": sample_bytes, "place_ms": scan_ms,
                "all_pairs_GBps": {k: sample_bytes / (v / 1e3) / 1e9 for k, v in scan_ms.items() if v > 0},
                "note": "the three placement kernels on the first segments of the workload, bit-identical results: the all-pairs "
                        "scan is what north_star describes (every read against every contig position; its rate IS a streamed-"
                        "bytes rate), the other two avoid the pairs"}

    # ---- strong scaling: the SAME 1000-segment study over the N ranks ----
    strong = None
    if world > 1 and not args.skip_strong:
        strong = strong_scaling(rig, batch, flags, ms_per_step)

    # the study's buffers are not needed any more: the contig-sharded configs want the memory
    del sb
    torch.cuda.empty_cache()

    sharded = None
    if sharded_plan is not None:
        import contig_sharded_bench as CS
        sharded = CS.run_gpu_side(me, rig, sharded_plan)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
            "reads_scored_per_s": all_reads / (ms_per_step / 1e3),
            "config": study_config(args.segments, world),
            "workload_stats": {"reads_per_gpu": N, "contigs_per_gpu": Cn, "cpu_affinity": rig.affinity},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms, "matches_device_resident_run": same,
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
        if strong:
            line["strong_scaling"] = strong
        if sharded:
            line["contig_sharded"] = sharded
        if cpu_base:
            line["cpu_baseline"] = cpu_base
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    sc.close()
    return 0


def strong_scaling(rig, batch0, flags, one_gpu_ms):
    """north_star: 'near-linear 1->8 GPU scaling on the 1k-segment study'.  The study of seeds SEED..SEED+S-1 (rank 0's
    weak-scaling batch) is cut into N contiguous runs of segments; every rank scores its run (device resident and
    host-to-host) and the records are all-gathered.  The one-GPU time beside it is the headline's (same study, same run)."""
    torch, dist, B, sc, args = rig.torch, rig.dist, rig.B, rig.sc, rig.args
    world, rank = rig.world, rig.rank
    S = args.segments
    bounds = [S * r // world for r in range(world + 1)]
    s0, s1 = bounds[rank], bounds[rank + 1]
    if rank == 0:
        part = [np.ascontiguousarray(x) if x is not None and not isinstance(x, int) else x for x in sharding.slice_batch(batch0, s0, s1)]
        mine = synth.Batch(*part)
    else:
        mine = synth.make_batch(s1 - s0, seed=SEED + s0, length=LENGTH, read_len=READ_LEN, coverage=COVERAGE)
    sb = StudyBuffers(rig, mine)
    nc = mine.n_contigs
    counts = rig.sum_over_ranks([float(nc if r == rank else 0) for r in range(world)])
    pad = int(max(counts))
    rec = torch.zeros(5, pad, dtype=torch.float64, device=rig.dev)
    out = torch.empty(world * 5 * pad, dtype=torch.float64, device=rig.dev)
    dflags = flags | B.DEVICE_CHARS | B.DEVICE_RESULT
    launches0 = sc.launch_count

    def step_device():
        sc.score_batch_raw(sb.db, sb.dr, 8, dflags)
        rec[:, :nc].copy_(sb.d_f64[:, :nc])
        dist.all_gather_into_tensor(out, rec.view(-1))

    dev_ms = rig.time_device(step_device, args.steps, args.warmup)
    launches = (sc.launch_count - launches0) // (args.steps + args.warmup)
    sb.make_host(rig)

    def step_host():
        sc.score_batch_raw(sb.hb, sb.hr, 8, flags)
        rec[:, :nc].copy_(sb.h_f64[:, :nc], non_blocking=True)
        dist.all_gather_into_tensor(out, rec.view(-1))

    host_ms = rig.time_wall(step_host, args.steps, 1)
    tot_pair, = rig.sum_over_ranks([mine.pair_bases()])
    gather_ms = rig.time_device(lambda: dist.all_gather_into_tensor(out, rec.view(-1)), 20, 3)  # the exchange alone
    res = None
    if rank == 0:
        res = {"segments_total": S, "segments_per_rank": [bounds[r + 1] - bounds[r] for r in range(world)], "scaling": "strong",
               "ms_per_step": dev_ms, "value": tot_pair / 1e9 / (dev_ms / 1e3), "unit": UNIT,
               "ms_per_step_one_gpu_same_run": one_gpu_ms,
               "speedup_vs_one_gpu_same_run": one_gpu_ms / dev_ms, "efficiency": one_gpu_ms / dev_ms / world,
               "e2e": {"ms_per_step": host_ms, "value": tot_pair / 1e9 / (host_ms / 1e3), "unit": UNIT},
               "fixed_costs": {"kernel_launches_per_step_per_rank": int(launches), "all_gather_of_records_ms": gather_ms,
                               "record_bytes_per_rank": int(5 * pad * 8)},
               "what": f"the {S}-segment study (seeds {SEED}..{SEED + S - 1}) split into {world} contiguous runs of segments; "
                       f"one bs_score_batch call per rank + one all_gather of the 5-column f64 record block; "
                       f"one-GPU time = rank 0's headline step (the same study, the gather-free N=1 path)"}
    del sb
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--segments", type=int, default=1000, help="segments per GPU (cfg-2 study size: 1000)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--device-only", action="store_true", help="device-resident phase only (profiling runs)")
    ap.add_argument("--no-study", action="store_true", help="skip the extra simulate-on-device / two-table measurements")
    ap.add_argument("--scan-segments", type=int, default=20, help="segments timed with the all-pairs / tile placement kernels (0: skip)")
    ap.add_argument("--skip-strong", action="store_true", help="skip the strong-scaling measurement of the study (N > 1)")
    ap.add_argument("--skip-sharded", action="store_true", help="skip the contig-sharded cfg-4 / cfg-5 measurements")
    ap.add_argument("--cfg5-scale", type=float, default=1.0, help="cfg-5 size as a fraction of 100 Mb / 1e8 reads / 1e5 contigs")
    ap.add_argument("--cfg4-scaffolds", type=int, default=10000)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        print("bench.py: warning: fewer than 3 warm-up steps", file=sys.stderr)
    if args.impl == "reference":
        return reference_main(args)
    return b200_main(args)


if __name__ == "__main__":
    sys.exit(main())
