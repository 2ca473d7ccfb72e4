import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def emulating() -> bool:
    """DANN_EMULATE=1: the `-m gpu` tests run on the CPU against tests/simt/_build/libdiskann_b200_emu.so - the
    product's C ABI (diskann_b200.cu + every kernel source) built by g++ on the SIMT emulator and a fake CUDA runtime.
    A logic check of host code and kernels together; see tests/test_emulated_abi.py.  Never set on the GPU box."""
    return os.environ.get("DANN_EMULATE") == "1"


def buffer_device():
    """torch device for the tests that hand device pointers to the C ABI ("device memory" is host memory when emulating)."""
    import torch
    return torch.device("cpu") if emulating() else torch.device("cuda", 0)


def dptr(t):
    """What a test hands to the ctypes mirror as a device buffer: the CUDA tensor itself, or - when emulating, where the
    mirror would (rightly) refuse a CPU tensor - the raw address of the host tensor standing in for device memory."""
    if not emulating() or t is None:
        return t

    class _HostAsDevice:                      # quacks like the CUDA tensor the mirror insists on
        is_cuda = True

        def __init__(self, x):
            self._x, self.shape = x, x.shape

        def is_contiguous(self):
            return self._x.is_contiguous()

        def data_ptr(self):
            return self._x.data_ptr()

        def numel(self):
            return self._x.numel()

    return _HostAsDevice(t)


@pytest.fixture(scope="session")
def lib_built():
    """The in-tree CUDA library, compiled if stale (nvcc cross-compiles without a GPU)."""
    if emulating():
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "simt"))
        import build_emu
        from pgvectorscale_b200 import diskann
        so = build_emu.build_abi(asan=os.environ.get("DANN_EMULATE_ASAN") == "1")   # ASan build: needs LD_PRELOAD=libasan
        diskann._LIB = None
        diskann._LIB = diskann.load_library(so)      # explicit path: the package itself never looks for this file
        return so
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
