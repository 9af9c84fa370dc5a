"""Reader of the fixtures written by baseline/dump_fixtures.jl (the real Julia reference, when one is available):
Int64 rank, Int64 dims..., Float64 data in column-major order (complex: re, im interleaved).

    arr = load("tests/golden/ref/c2_segments.bin")            # (11, 8, n_rays), Fortran order
    fld = load("tests/golden/ref/c1_field.bin", complex=True)

tests/test_golden.py picks the files up when tests/golden/ref/ exists and compares the oracle AND the GPU path with them
(hit points / directions 1e-9 relative, field 1e-8 relative L2); without them the oracle-generated goldens stay in charge."""
import os

import numpy as np


def load(path, complex=False):
    raw = np.fromfile(path, dtype=np.int64, count=1)
    rank = int(raw[0])
    dims = np.fromfile(path, dtype=np.int64, count=1 + rank)[1:]
    data = np.fromfile(path, dtype=np.float64, offset=8 * (1 + rank))
    if complex:
        data = data[0::2] + 1j * data[1::2]
    return data.reshape(tuple(int(d) for d in dims), order="F")


def available(root=None):
    root = root or os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref")
    return os.path.isdir(root) and os.path.exists(os.path.join(root, "c2_segments.bin"))
