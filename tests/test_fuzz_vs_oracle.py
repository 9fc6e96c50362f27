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
