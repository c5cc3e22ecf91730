"""End-to-end (host pointers, pinned) query timing at the headline config for the current NBK_HOST_SLICE."""
import os, sys, time
import torch
sys.path.insert(0, ".")
from nbodyhpc_b200 import capi
n, m, k = 512 ** 3, 100_000_000, 8
g = torch.Generator(device="cuda"); g.manual_seed(42)
pts = torch.rand((n, 3), device="cuda", generator=g)
tree = capi.Tree.build_device(pts.data_ptr(), n, 64, 1.0, stream=torch.cuda.current_stream().cuda_stream)
del pts
g.manual_seed(43)
q = torch.rand((m, 3), device="cuda", generator=g)
qh = torch.empty((m, 3), dtype=torch.float32, pin_memory=True); qh.copy_(q); del q
od = torch.empty((m, k), dtype=torch.float32, pin_memory=True)
oi = torch.empty((m, k), dtype=torch.int32, pin_memory=True)
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    tree.query_raw(qh.data_ptr(), m, k, od.data_ptr(), oi.data_ptr())
    dt = time.perf_counter() - t0
    print(os.environ.get("NBK_HOST_SLICE", "default"), f"iter {it}: {dt*1e3:.1f} ms  {m/dt/1e6:.1f} Mq/s", flush=True)
