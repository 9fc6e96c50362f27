"""cfg-4 scaffold set scored from its parts: stage times under a few settings (run on the GPU box)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from genomeassembler_dev_b200 import breakscore as B, synth, tables

kmers = tables.all_kmer_strings(); prob = tables.normalised(tables.load_raw())
seg = synth.make_scaffold_set(1400, n_scaffolds=int(os.environ.get("N_SCAF", "10000")))
sset = B.ScaffoldSet(seg.base_contigs, seg.part_start, seg.part_base, np.zeros(len(seg.part_base), np.int32))
sc = B.BreakageScorer(0)
sc.set_table(kmers, prob)
variants = [("default", {}, B.WANT_KS | B.WANT_STARTPOS), ("weights_path", {"BS_COMPOSE_SCORE": "0"}, B.WANT_KS | B.WANT_STARTPOS),
            ("no_ks", {}, B.WANT_STARTPOS), ("ks_a_only_via_rows_global", {"BS_COMPOSE_ROWS": "global"}, B.WANT_KS | B.WANT_STARTPOS),
            ("scores_only", {}, 0)]
for name, env, flags in variants:
    for k in ("BS_COMPOSE_SCORE", "BS_COMPOSE_ROWS"):
        os.environ.pop(k, None)
    os.environ.update(env)
    for _ in range(3):
        sc.score_scaffolds.__func__  # noqa
        lens = sset.lengths()
        res = sc.score_batch(seg.reads.reshape(-1), None, seg.reads.shape[1], np.zeros(1, np.uint8), np.concatenate([[0], np.cumsum(lens)]),
                             np.frombuffer(seg.truth, np.uint8), [0, len(seg.truth)], [0, len(seg.reads)], [0, len(sset)], flags=flags, scaffolds=sset)
    sc.enable_timing(True)
    t0 = time.perf_counter()
    res = sc.score_batch(seg.reads.reshape(-1), None, seg.reads.shape[1], np.zeros(1, np.uint8), np.concatenate([[0], np.cumsum(lens)]),
                         np.frombuffer(seg.truth, np.uint8), [0, len(seg.truth)], [0, len(seg.reads)], [0, len(sset)], flags=flags, scaffolds=sset)
    wall = 1e3 * (time.perf_counter() - t0)
    tm = {k: round(v, 3) for k, v in sc.last_timings().items() if v > 0}
    sc.enable_timing(False)
    print(json.dumps({"variant": name, "wall_ms": round(wall, 2), "stages": tm, "breaks": int(res["kmer_breaks"].sum())}), flush=True)
