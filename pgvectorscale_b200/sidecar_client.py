"""Client side of sidecar/dann_sidecar.c's wire protocol (what the pgrx shim's amrescan would speak, here in Python
for tests and tools).  One connection = one backend; scan() blocks until the sidecar's coalescer has run the query."""
from __future__ import annotations

import socket
import struct
from typing import Optional, Sequence

import numpy as np


class SidecarClient:
    def __init__(self, path: str, timeout: float = 60.0):
        self._s = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        self._s.settimeout(timeout)
        self._s.connect(path)
        magic, self.dim, self.n = struct.unpack("<III", self._recv(12))
        if magic != 0x484E4144:
            raise RuntimeError("not a dann_sidecar socket")

    def _recv(self, n: int) -> bytes:
        buf = bytearray()
        while len(buf) < n:
            chunk = self._s.recv(n - len(buf))
            if not chunk:
                raise ConnectionError("sidecar closed the connection")
            buf += chunk
        return bytes(buf)

    def scan(self, query, labels: Optional[Sequence[int]] = None, k: int = 10, search_list_size: int = 100,
             rescore: int = 50) -> dict:
        q = np.ascontiguousarray(query, dtype=np.float32)
        if q.shape != (self.dim,):
            raise ValueError(f"query must have {self.dim} dimensions")
        lab = np.asarray([] if labels is None else list(labels), dtype=np.int16)
        nl = -1 if labels is None else len(lab)
        self._s.sendall(struct.pack("<Iiiii", 0x514E4144, k, search_list_size, rescore, nl) + q.tobytes() + lab.tobytes())
        status, count = struct.unpack("<iI", self._recv(8))
        if status != 0:
            (ln,) = struct.unpack("<I", self._recv(4))
            raise RuntimeError(f"sidecar error {status}: {self._recv(ln).decode('utf-8', 'replace')}")
        tid = np.frombuffer(self._recv(8 * k), dtype=np.uint64).copy()
        dist = np.frombuffer(self._recv(4 * k), dtype=np.float32).copy()
        st = struct.unpack("<6I", self._recv(24))
        return dict(tid=tid, dist=dist, count=int(count),
                    stats=dict(zip(("visits", "d_quantized", "candidates", "d_full", "stream_len", "status"), st)))

    def close(self):
        self._s.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
