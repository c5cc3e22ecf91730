"""Kernel A/B on one GPU: device-resident query step of the headline workload at several batch sizes.

    [NBK_LIBRARY=nbodyhpc_b200/lib/variants/libnbk_X.so] python scripts/kernel_ab.py [--queries 100000000,12500000] [-k 8]

Prints one JSON line per batch size: ms per step (CUDA events), the library's own kNN / ordering section
times, and a checksum of the rows (equal checksums = identical results across variants)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nbodyhpc_b200 import capi  # noqa: E402
from scripts.synthetic import uniform, zeldovich  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--queries", default="100000000,12500000")
ap.add_argument("-k", type=int, default=8)
ap.add_argument("--side", type=int, default=512)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--clustered", action="store_true")
ap.add_argument("--open", action="store_true")
args = ap.parse_args()
dev = torch.device("cuda", 0)
n = args.side ** 3
pts = zeldovich(args.side, 42, dev) if args.clustered else uniform(n, 42, dev)
stream = torch.cuda.current_stream().cuda_stream
tree = capi.Tree.build_device(pts.data_ptr(), n, 64, None if args.open else 1.0, stream=stream)
del pts
for m in [int(x) for x in args.queries.split(",")]:
    q = uniform(m, 43, dev)
    d = torch.empty((m, args.k), device=dev)
    i = torch.empty((m, args.k), device=dev, dtype=torch.int32)
    for _ in range(2):
        tree.query_device(q.data_ptr(), m, args.k, d.data_ptr(), i.data_ptr(), stream)
    torch.cuda.synchronize()
    capi.profile_read(capi.SECTION_KNN_KERNEL); capi.profile_read(capi.SECTION_QUERY_ORDER)
    capi.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        tree.query_device(q.data_ptr(), m, args.k, d.data_ptr(), i.data_ptr(), stream)
    e1.record(); torch.cuda.synchronize()
    capi.profile_enable(False)
    knn_ms, _ = capi.profile_read(capi.SECTION_KNN_KERNEL)
    order_ms, _ = capi.profile_read(capi.SECTION_QUERY_ORDER)
    ms = e0.elapsed_time(e1) / args.steps
    print(json.dumps({"lib": os.path.basename(capi.library_path()), "env": {k: v for k, v in os.environ.items() if k.startswith("NBK_") and k != "NBK_LIBRARY"},
                      "m": m, "k": args.k, "ms_per_step": round(ms, 3), "knn_ms": round(knn_ms / args.steps, 3),
                      "order_ms": round(order_ms / args.steps, 3), "gqps": round(m / ms / 1e6, 4),
                      "checksum": [int(i.long().sum()), int(d.view(torch.int32).long().sum())]}), flush=True)
    del q, d, i
