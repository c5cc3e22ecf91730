"""End-to-end timing of nbk_tree_query with pinned host buffers (what bench.py's `e2e` leg measures), for tuning
the host pipeline: NBK_HOST_SLICE / NBK_HOST_FIRST_SLICE.

    python scripts/e2e_probe.py [queries] [k]"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nbodyhpc_b200 import capi  # noqa: E402
from scripts.synthetic import uniform  # noqa: E402

m = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device("cuda", 0)
pts = uniform(512 ** 3, 42, dev)
tree = capi.Tree.build_device(pts.data_ptr(), 512 ** 3, 64, 1.0, stream=torch.cuda.current_stream().cuda_stream)
del pts
q = torch.empty((m, 3), dtype=torch.float32, pin_memory=True)
q.copy_(uniform(m, 43, dev))
d = torch.empty((m, k), dtype=torch.float32, pin_memory=True)
i = torch.empty((m, k), dtype=torch.int32, pin_memory=True)
tree.query_raw(q.data_ptr(), m, k, d.data_ptr(), i.data_ptr())
times = []
for _ in range(4):
    t0 = time.perf_counter()
    tree.query_raw(q.data_ptr(), m, k, d.data_ptr(), i.data_ptr())
    times.append(time.perf_counter() - t0)
print(json.dumps({"env": {a: b for a, b in os.environ.items() if a.startswith("NBK_HOST")}, "m": m, "k": k,
                  "seconds": [round(t, 4) for t in times], "gqps_best": round(m / min(times) / 1e9, 4)}), flush=True)
