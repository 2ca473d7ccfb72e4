import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def lib_built():
    """The in-tree CUDA library, compiled if stale (nvcc cross-compiles without a GPU)."""
    from pgvectorscale_b200.build import build_library
    return build_library()


def build_case(n, dim, distance_type, bits=None, seed=1, kind="normal", R=50, L_build=100,
               labels=False, dim_index=None, deleted_every=0, train_on_data=True):
    """Index fixture built through the oracle (train -> quantize -> serial Vamana)."""
    from oracle import fixtures
    v = fixtures.gen_vectors(n, dim, seed, kind)
    label_off = lab = None
    if labels:
        label_off, lab = fixtures.gen_labels(n, seed + 1000)
    s = fixtures.make_index(v, distance_type, bits=bits, R=R, L_build=L_build, dim_index=dim_index,
                            label_off=label_off, labels=lab, train_on_data=train_on_data)
    if deleted_every:
        # vacuum marks a node deleted by invalidating its heap offset (vacuum.rs, scan.rs:231-234)
        t = s.heap_tid.copy()
        t[::deleted_every] &= np.uint64(0xFFFFFFFFFFFF0000)
        s.heap_tid = t
    return s
