"""B200-native breakage scorer: the hot path of SahakyanLab/GenomeAssembler_dev
(``calc_breakscore`` of ``lib/BreakageScorer.cpp``) behind the reference's own interface.

* :mod:`.breakscore` -- ctypes mirror of the C-ABI (``include/breakscore.h``): ``calc_breakscore``,
  ``assemble_contigs``, :class:`BreakageScorer` (``score``, ``score_batch``, ``simulate_reads``).
* :mod:`.tables` -- the breakage-probability tables (``bp_kmer`` / ``bp_prob``).
* :mod:`.synth` -- seeded synthetic workloads of the BASELINE.json shapes (host-side input generation).
* :mod:`.sharding` -- partitioning of a job over the GPUs of one box and the record gather.

Everything is computed by ``libbreakscore.so`` (CUDA, sm_100a); there is no CPU fallback.
"""
from .breakscore import BreakageScorer, BreakscoreError, assemble_contigs, calc_breakscore  # noqa: F401

__all__ = ["BreakageScorer", "BreakscoreError", "assemble_contigs", "calc_breakscore"]
