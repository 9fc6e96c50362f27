#!/usr/bin/env python
"""Stage timings of the scorer on the other BASELINE.json shapes (not the bench line): cfg-1 single
segment, cfg-3 read-length sweep, cfg-4 scaffold set, cfg-5 scaled.  Prints one JSON line per shape."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genomeassembler_dev_b200 import breakscore as B, synth, tables  # noqa: E402


def run(sc, name, seg, reps=3, flags=B.DEFAULT_FLAGS):
    ct, ct_off = B.flatten(seg.contigs)
    rd = np.ascontiguousarray(seg.reads).reshape(-1)
    tr, tr_off = B.flatten([seg.truth])
    args = (rd, None, seg.reads.shape[1], ct, ct_off, tr, tr_off, [0, seg.reads.shape[0]], [0, len(seg.contigs)])
    for _ in range(3):  # every one of the three rotating workspaces grows to this shape before the clock starts
        sc.score_batch(*args, flags=flags)
    sc.enable_timing(True)
    t0 = time.perf_counter()
    for _ in range(reps):
        res = sc.score_batch(*args, flags=flags)
    wall = (time.perf_counter() - t0) / reps
    st = {k: round(v / reps, 3) for k, v in sc.last_timings().items() if v >= 0}
    sc.enable_timing(False)
    pair = seg.reads.shape[0] * float(sum(len(c) for c in seg.contigs))
    print(json.dumps({"shape": name, "reads": int(seg.reads.shape[0]), "contigs": len(seg.contigs),
                      "contig_bases": int(sum(len(c) for c in seg.contigs)), "wall_ms": round(wall * 1e3, 2),
                      "pair_Gbp_per_s_wall": round(pair / 1e9 / wall, 1), "stage_ms": st,
                      "placed": int(res["kmer_breaks"].sum())}), flush=True)


def main():
    kmers, prob = tables.all_kmer_strings(), tables.normalised(tables.load_raw())
    sc = B.BreakageScorer(0)
    sc.set_table(kmers, prob)
    run(sc, "cfg1: 50 kb, r=100 30x, 16 contigs", synth.make_segment(1234, 50000, 100, 30, 16))
    run(sc, "cfg1 + edit distance (BS_WANT_LEV)", synth.make_segment(1234, 50000, 100, 30, 16), flags=B.DEFAULT_FLAGS | B.WANT_LEV)
    b = synth.make_batch(200, seed=1234)
    class _S:  # 200 cfg-2 segments with the edit distance on, through score_batch directly
        pass
    sc.enable_timing(True)
    args = (b.read_chars, None, b.read_len, b.contig_chars, b.contig_off, b.truth_chars, b.truth_off, b.seg_read_start, b.seg_contig_start)
    for fl, nm in ((B.DEFAULT_FLAGS, "cfg2 x200 segments"), (B.DEFAULT_FLAGS | B.WANT_LEV, "cfg2 x200 segments + edit distance")):
        sc.score_batch(*args, flags=fl)
        sc.last_timings()
        t0 = time.perf_counter()
        res = sc.score_batch(*args, flags=fl)
        wall = time.perf_counter() - t0
        print(json.dumps({"shape": nm, "contigs": b.n_contigs, "wall_ms": round(wall * 1e3, 2),
                          "stage_ms": {k: round(v, 3) for k, v in sc.last_timings().items() if v >= 0},
                          "lev_hist": np.bincount(np.minimum(res["lev_dist_vs_true"], 3)).tolist()}), flush=True)
    sc.enable_timing(False)
    for r in (50, 300):
        run(sc, f"cfg3: 50 kb, r={r} 30x, 40 contigs", synth.make_segment(1300 + r, 50000, r, 30, 40))
    run(sc, "cfg3: 50 kb, r=12 40x, 60 contigs (script 00 grid)", synth.make_segment(1312, 50000, 12, 40, 60))
    cfg4 = synth.make_scaffold_set(1400, n_scaffolds=10000)
    run(sc, "cfg4: 10000 scaffolds of 10-50 kb, one segment", cfg4, reps=2, flags=B.WANT_KS | B.WANT_STARTPOS)
    # SHAPE_MULTI_GPUS=N: the same scaffold set with its contigs dealt out over N GPUs of this process (bs_score_multi)
    n_gpus = int(os.environ.get("SHAPE_MULTI_GPUS", "1"))
    if n_gpus > 1:
        others = [B.BreakageScorer(d) for d in range(1, n_gpus)]
        for o in others:
            o.set_table(kmers, prob)
        fl = B.WANT_KS | B.WANT_STARTPOS
        one = sc.score(cfg4.contigs, cfg4.reads, cfg4.truth, flags=fl)
        for _ in range(2):
            sc.score(cfg4.contigs, cfg4.reads, cfg4.truth, flags=fl, group=others)
        t0 = time.perf_counter()
        got = sc.score(cfg4.contigs, cfg4.reads, cfg4.truth, flags=fl, group=others)
        wall = time.perf_counter() - t0
        same = all(np.array_equal(one[k], got[k], equal_nan=True) for k in one if k != "sequence")
        pair = cfg4.reads.shape[0] * float(sum(len(c) for c in cfg4.contigs))
        print(json.dumps({"shape": f"cfg4 over {n_gpus} GPUs of one process (bs_score_multi)", "wall_ms": round(wall * 1e3, 2),
                          "pair_Gbp_per_s_wall": round(pair / 1e9 / wall, 1), "identical_to_one_gpu": bool(same)}), flush=True)
        for o in others:
            o.close()
    rng = np.random.default_rng(1500)
    L, N, Cn, r = 10_000_000, 2_000_000, 10_000, 150
    truth = synth.codes_to_ascii(synth.random_truth_codes(rng, L))
    starts = rng.integers(0, L - r, size=N)
    reads = truth[starts[:, None] + np.arange(r)[None, :]]
    cstart = np.sort(rng.integers(0, L - 3000, size=Cn))
    clen = rng.integers(200, 2000, size=Cn)
    seg = synth.Segment(truth.tobytes(), reads, [truth[a:a + b].tobytes() for a, b in zip(cstart, clen)])
    # 888 resident blocks x 2e6 reads x 8 B of dense rows exceed the 4 GB scratch budget: hashed scratch by default
    run(sc, "cfg5 scaled 1/10: 10 Mb truth, 2e6 reads, 1e4 contigs of ~1 kb (hashed placement scratch)", seg, reps=2)
    os.environ["BS_PLACE_SCRATCH_MB"] = "65536"
    run(sc, "cfg5 scaled 1/10, dense placement scratch (14 GB of rows)", seg, reps=2)
    del os.environ["BS_PLACE_SCRATCH_MB"]


if __name__ == "__main__":
    main()
