"""Throughput vs k at the headline tree (512^3 uniform periodic, 1e8 queries)."""
import json, sys, torch
sys.path.insert(0, ".")
from nbodyhpc_b200 import capi
from scripts.config_sweep import uniform, check
side = int(sys.argv[1]) if len(sys.argv) > 1 else 512
ks = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [16, 32, 64]
n, m = side ** 3, 100_000_000
pts, q = uniform(n, 42), uniform(m, 43)
stream = torch.cuda.current_stream().cuda_stream
tree = capi.Tree.build_device(pts.data_ptr(), n, 64, 1.0, stream=stream)
for k in ks:
    od = torch.empty((m, k), device="cuda", dtype=torch.float32); oi = torch.empty((m, k), device="cuda", dtype=torch.int32)
    ts = []
    for _ in range(2):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); tree.query_device(q.data_ptr(), m, k, od.data_ptr(), oi.data_ptr(), stream); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    row = {"side": side, "k": k, "ms": min(ts), "mq_per_s": m / min(ts) / 1e3}
    if "--check" in sys.argv: row["check"] = check(pts, q, od, oi, k, 1.0, n, sample=64)
    print(json.dumps(row), flush=True)
    del od, oi
