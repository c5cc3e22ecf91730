"""Small end-to-end pass over every kernel family (build top + bottom phases, radix fallback, ordering,
register / shared-memory / global-memory top-k containers, boundary pass, CDF, statistics, flat-block scan):
a quick "does every kernel run" check, e.g. under a memory checker where one is available.

    python scripts/all_kernels_probe.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nbodyhpc_b200 import capi  # noqa: E402

rng = np.random.default_rng(0)
pts = rng.random((30_011, 3), dtype=np.float32)
q = rng.random((5_003, 3), dtype=np.float32)
q[:2000] = np.where(rng.random((2000, 3)) < 0.5, q[:2000] * 2e-3, 1 - q[:2000] * 2e-3).astype(np.float32)  # face huggers
edges = np.linspace(0, 0.3, 17).astype(np.float32)
for box in (None, 1.0):
    tree = capi.Tree.build(pts, 32, box)
    tree.nodes(); tree.points()
    for k in (1, 3, 8, 16, 33, 70):
        d, i = tree.query(q, k, squared=(k % 2 == 0))
        assert (np.diff(d, axis=1) >= 0).all()
    tree.knn_cdf(q, [1, 8, 40, 70], edges)
    tree.stats(q[:500], 8); tree.stats(q[:200], 70)
    tree.close()
capi.Tree.build(np.full((5000, 3), 0.5, np.float32), 16).close()  # ties: the radix fallback of the bottom kernel
x = pts[:256]
capi.scan_block(x[:, 0], x[:, 1], x[:, 2], np.arange(256, dtype=np.uint32), q[:64], 20, boxsize=1.1)
big = rng.random((600_000, 3), dtype=np.float32)  # several ordering tiles, top-phase levels
t = capi.Tree.build(big, 64, 1.0)
t.query(rng.random((300_000, 3), dtype=np.float32), 8)
print("all kernel families ran")
