#!/bin/bash
# which commit made the GPU builder leave nodes unconnected (avg_degree 50.0 -> 48.9 at 1M, recall plateau 0.94)?
# every candidate library builds the 1M fixture; the in-tree library is restored afterwards
mkdir -p gpurun_out
cp pgvectorscale_b200/libdiskann_b200.so /tmp/lib_head.so
: > gpurun_out/r2x_bisect.txt
for so in _bisect/lib_71151ab.so _bisect/lib_950515b.so _bisect/lib_b224633.so _bisect/lib_e0113f2.so _bisect/lib_1b38bc0.so /tmp/lib_head.so; do
  [ -f "$so" ] || continue
  cp "$so" pgvectorscale_b200/libdiskann_b200.so
  echo "== $so" >> gpurun_out/r2x_bisect.txt
  timeout 300 python - >> gpurun_out/r2x_bisect.txt 2>&1 <<'PY'
import sys, torch
sys.path.insert(0, '.')
from tools import fixture as fx
import numpy as np
dev = torch.device('cuda', 0)
snap, idx, st = fx.codes_and_graph(1_000_000, 768, 'lowrank', 0, dev, log=lambda *a: None, download_nbrs=True)
nb = snap.nbrs
deg = (nb != 0xFFFFFFFF).sum(1)
first_inv = np.argmax(nb == 0xFFFFFFFF, axis=1)
holes = int(((nb == 0xFFFFFFFF).any(1) & (deg > first_inv)).sum())
indeg = np.bincount(nb[nb != 0xFFFFFFFF].astype(np.int64), minlength=snap.n)
print('avg_degree', st['avg_degree'], 'deg0', int((deg == 0).sum()), 'deg<50', int((deg < 50).sum()), 'lists with a hole', holes,
      'no in-edge', int((indeg == 0).sum()), 'deg hist low', np.bincount(deg, minlength=65)[:8].tolist())
z = np.nonzero(deg < 50)[0]
print('first short nodes', z[:12].tolist(), 'last', z[-5:].tolist() if len(z) else [])
idx.close()
PY
done
cp /tmp/lib_head.so pgvectorscale_b200/libdiskann_b200.so
cat gpurun_out/r2x_bisect.txt
