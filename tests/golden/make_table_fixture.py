"""Regenerates tests/golden/query_table_k2468.npz from the upstream data files.

Run in the build container (where /root/reference exists):
    python tests/golden/make_table_fixture.py
The GPU box has no /root/reference, so the table -- input DATA of the hot path, upstream
data/QueryTable/QueryTable_kmer-{2,4,6,8}.csv, 69 904 rows -- is committed as a fixture.
Row order (2-,4-,6-,8-mers, each lexicographic ACGT) is asserted while reading.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from genomeassembler_dev_b200 import tables  # noqa: E402

raw = tables.load_raw_from_csv("/root/reference/data/QueryTable")
assert raw.shape == (tables.N_ROWS,)
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "query_table_k2468.npz")
np.savez_compressed(out, raw_prob=raw)
print(out, raw.shape, raw.sum(), raw.min(), raw.max())
