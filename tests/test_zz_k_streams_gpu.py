"""Device-buffer entry points called with stream = NULL run on the index's own non-blocking stream, ordered after
whatever the caller has already submitted to the legacy default stream (include/diskann_b200.h, "entry points that take
DEVICE buffers").  Round 2 found the hole on B200: the bench fixture quantized dataset chunks that torch was still
generating on its default stream - the GPU builder then left a fifth of the nodes of a 1M index with short lists and the
recall of every scan plateaued at 0.94 - because the blocking cudaMemcpy calls of the index load, which had been the
accidental barrier, had become stream-ordered copies."""
import numpy as np
import pytest
import torch

from conftest import build_case, buffer_device, dptr, emulating
from pgvectorscale_b200.snapshot import COSINE

pytestmark = pytest.mark.gpu


def test_null_stream_calls_wait_for_the_callers_default_stream_work(lib_built):
    from pgvectorscale_b200 import diskann
    dev = buffer_device()
    s = build_case(600, 768, COSINE, seed=3, R=16, L_build=32)
    B = 4096
    with diskann.DiskAnnIndex(s) as idx:
        cw = idx.code_stride
        g = torch.Generator(device=dev)
        g.manual_seed(5)
        base = torch.randn((B, 768), generator=g, device=dev)
        big = torch.randn((4096, 4096), generator=g, device=dev)
        want = None
        for attempt in range(3):
            if not emulating():
                torch.cuda.synchronize()
                y = big
                for _ in range(40):                      # tens of milliseconds of default-stream work in front of x
                    y = torch.tanh(y @ big * 1e-2)
            else:
                y = big
            x = (base + y[:B, :768] * 0.0 + float(attempt)).contiguous()      # x is ready only when the chain above is
            codes = torch.zeros((B, cw), dtype=torch.int64, device=dev)
            qfull = torch.zeros((B, 768), dtype=torch.float32, device=dev)
            idx.prepare_queries(dptr(x), dptr(qfull), dptr(codes))          # stream = NULL, no synchronisation by the caller
            got = codes.cpu().numpy().copy()
            if not emulating():
                torch.cuda.synchronize()
            ref = torch.zeros((B, cw), dtype=torch.int64, device=dev)
            idx.prepare_queries(dptr(x), dptr(qfull), dptr(ref))            # the same call on inputs that are certainly complete
            assert np.array_equal(got, ref.cpu().numpy()), f"attempt {attempt}: codes of rows that were still being produced"
