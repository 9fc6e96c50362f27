"""TEST INFRASTRUCTURE ONLY: CPU oracle for the breakage-scoring hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
this package.  The product package ``genomeassembler_dev_b200`` never does.
"""
