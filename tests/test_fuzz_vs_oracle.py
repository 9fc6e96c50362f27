"""Randomised differential tests (hypothesis): small adversarial inputs -- tiny alphabets, repeats,
N runs, lowercase, ragged read lengths, contigs shorter than the seed, reads longer than contigs --
scored by the device algorithm and by the oracle.  CPU: the kernel sources under the emulation;
GPU (-m gpu): the nvcc build."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import parity_cases as P
from genomeassembler_dev_b200 import breakscore as B
from genomeassembler_dev_b200.synth import Segment

ALPHABETS = ["ACGT", "AC", "A", "ACGTN", "ACGTacgtN", "AT"]


@st.composite
def scoring_problem(draw):
    alpha = draw(st.sampled_from(ALPHABETS))
    seq = st.text(alphabet=alpha, min_size=0, max_size=120)
    truth = draw(st.text(alphabet=alpha, min_size=1, max_size=400))
    n_ctg = draw(st.integers(1, 5))
    contigs = []
    for _ in range(n_ctg):
        if draw(st.booleans()) and len(truth) > 2:  # a substring of the truth, possibly mutated
            a = draw(st.integers(0, len(truth) - 1))
            b = draw(st.integers(a, len(truth)))
            c = truth[a:b]
            if c and draw(st.booleans()):
                i = draw(st.integers(0, len(c) - 1))
                c = c[:i] + draw(st.sampled_from(alpha)) + c[i + 1:]
            contigs.append(c)
        else:
            contigs.append(draw(seq))
    uniform = draw(st.booleans())
    rl = draw(st.integers(1, 70))
    reads = []
    for _ in range(draw(st.integers(0, 40))):
        src = draw(st.sampled_from(contigs + [truth]))
        ln = rl if uniform else draw(st.integers(0, 70))
        if len(src) >= ln and draw(st.integers(0, 9)) < 8:
            a = draw(st.integers(0, len(src) - ln))
            reads.append(src[a:a + ln])
        else:
            reads.append(draw(st.text(alphabet=alpha, min_size=ln, max_size=ln)))
    kmer = draw(st.sampled_from([8, 8, 8, 4, 6, 2]))
    mode = draw(st.sampled_from([0, 0, B.PLACE_TILE, B.PLACE_SCAN]))
    return contigs, reads, truth, kmer, mode


def run(scorer, oracle, kmers, prob, problem):
    contigs, reads, truth, kmer, mode = problem
    seg = Segment(truth.encode(), None, [c.encode() for c in contigs])
    P.check_segment(scorer, oracle, kmers, prob, seg, kmer=kmer, reads=[r.encode() for r in reads], flags=P.FULL | mode)


@settings(max_examples=60, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(problem=scoring_problem())
def test_fuzz_emulated_device_algorithm(problem, emul_scorer, oracle, kmers, prob):
    run(emul_scorer, oracle, kmers, prob, problem)


@pytest.mark.gpu
@settings(max_examples=300, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(problem=scoring_problem())
def test_fuzz_gpu(problem, gpu_scorer, oracle, kmers, prob):
    run(gpu_scorer, oracle, kmers, prob, problem)


# ---- scaffold sets given as parts (bs_score_scaffolds): random valid sets, scored from the parts vs the oracle on the texts ----

@st.composite
def scaffold_problem(draw):
    alpha = draw(st.sampled_from(["ACGT", "AC", "A", "ACGTN", "AT"]))
    genome = draw(st.text(alphabet=alpha, min_size=20, max_size=300))
    n_base = draw(st.integers(1, 6))
    base = []
    for _ in range(n_base):
        a = draw(st.integers(0, len(genome) - 1))
        b = draw(st.integers(a + 1, min(len(genome), a + 90)))
        base.append(genome[a:b])
    chains = []
    for _ in range(draw(st.integers(1, 6))):
        text, chain, last_start = "", [], 0
        for _ in range(draw(st.integers(1, 5))):
            b = draw(st.integers(0, n_base - 1))
            # overlaps that are true suffix/prefix matches, that leave the part a base to add and keep the starts ascending
            valid = [k for k in range(0, min(len(text), len(base[b]) - 1) + 1)
                     if text.endswith(base[b][:k]) and len(text) - k >= last_start]
            if not valid:
                continue
            k = draw(st.sampled_from(valid)) if draw(st.booleans()) else max(valid)
            chain.append((b, k))
            last_start = len(text) - k
            text += base[b][k:]
        if chain:
            chains.append(chain)
    if not chains:
        chains = [[(0, 0)]]
    uniform = draw(st.booleans())
    rl = draw(st.integers(1, 40))
    texts = ["".join(base[b][k:] for b, k in ch) for ch in chains]
    reads = []
    for _ in range(draw(st.integers(0, 30))):
        src = draw(st.sampled_from(texts + [genome]))
        ln = rl if uniform else draw(st.integers(0, 40))
        if len(src) >= ln and draw(st.integers(0, 9)) < 8:
            a = draw(st.integers(0, len(src) - ln))
            reads.append(src[a:a + ln])
        else:
            reads.append(draw(st.text(alphabet=alpha, min_size=ln, max_size=ln)))
    kmer = draw(st.sampled_from([8, 8, 8, 4, 6]))
    mode = draw(st.sampled_from(["scored_in_place", "scored_in_place", "weights", "ks_from_parts", "junctions_probed", "small_hash"]))
    return base, chains, reads, genome, kmer, mode


MODE_ENV = {"weights": ("BS_COMPOSE_SCORE", "0"), "junctions_probed": ("BS_COMPOSE_JUNCTIONS", "0"), "small_hash": ("BS_COMPOSE_HASH_SLOTS", "64")}


def run_scaffolds(scorer, oracle, kmers, prob, problem, lib_path=None):
    import os
    import scaffold_cases as SC
    base, chains, reads, genome, kmer, mode = problem
    sset = SC.hand_scaffold_set([b.encode() for b in base], chains, lib_path)
    assert sset.texts() == ["".join(base[b][k:] for b, k in ch).encode() for ch in chains]
    env = MODE_ENV.get(mode)
    if env:
        os.environ[env[0]] = env[1]
    try:
        SC.check_scaffolds(scorer, oracle, kmers, prob, genome.encode(), [r.encode() for r in reads], sset, kmer=kmer, flags=SC.mode_flags(mode))
    finally:
        if env:
            del os.environ[env[0]]


@settings(max_examples=50, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(problem=scaffold_problem())
def test_fuzz_scaffold_sets_emulated(problem, emul_scorer, emul_lib, oracle, kmers, prob):
    run_scaffolds(emul_scorer, oracle, kmers, prob, problem, emul_lib)


@pytest.mark.gpu
@settings(max_examples=300, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(problem=scaffold_problem())
def test_fuzz_scaffold_sets_gpu(problem, gpu_scorer, oracle, kmers, prob):
    run_scaffolds(gpu_scorer, oracle, kmers, prob, problem)
