"""Quick on-GPU probe: build + query timings at a few scales, parity spot checks at scale."""
import json, sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from nbodyhpc_b200 import capi

def ev():
    return torch.cuda.Event(enable_timing=True)

def run(n_side, m, k, leaf, box=1.0, reps=3, check=False):
    n = n_side ** 3
    g = torch.Generator(device="cuda"); g.manual_seed(42)
    pts = torch.rand((n, 3), device="cuda", generator=g)
    g.manual_seed(43)
    q = torch.rand((m, 3), device="cuda", generator=g)
    stream = torch.cuda.current_stream().cuda_stream
    torch.cuda.synchronize()
    t0 = time.time()
    tree = capi.Tree.build_device(pts.data_ptr(), n, leaf, box, stream=stream)
    torch.cuda.synchronize()
    t_build = time.time() - t0
    t0 = time.time()
    tree2 = capi.Tree.build_device(pts.data_ptr(), n, leaf, box, stream=stream)
    torch.cuda.synchronize()
    t_build2 = time.time() - t0
    tree2.close()
    od = torch.empty((m, k), device="cuda", dtype=torch.float32)
    oi = torch.empty((m, k), device="cuda", dtype=torch.int32)
    times = []
    for r in range(reps):
        a, b = ev(), ev()
        a.record()
        tree.query_device(q.data_ptr(), m, k, od.data_ptr(), oi.data_ptr(), stream)
        b.record(); torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    res = dict(n=n, m=m, k=k, leaf=leaf, box=box, build_s=t_build, build2_s=t_build2,
               build_mpts=n / t_build2 / 1e6, query_ms=times, mqps=m / min(times) / 1e3,
               nodes=tree.size, free_gb=torch.cuda.mem_get_info()[0] / 2**30)
    if check:
        from oracle import Reference, compare_knn
        ms = min(m, 20000)
        hp = pts.cpu().numpy(); hq = q[:ms].cpu().numpy()
        ref = Reference.Tree(hp, leaf, box)
        d_ref, i_ref, st = ref.query(hq, k, workers=0, return_stats=True)
        rep = compare_knn(od[:ms].cpu().numpy(), oi[:ms].cpu().numpy().view(np.uint32), d_ref, i_ref, hp, hq, box)
        res["parity"] = str(rep); res["ref_stats_per_query"] = (st / ms).tolist()
        res["gpu_stats_per_query"] = (tree.stats(hq, k) / ms).tolist()
    print(json.dumps(res), flush=True)
    tree.close()

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), torch.cuda.mem_get_info())
    run(100, 1_000_000, 8, 64, check=True)
    run(128, 2_097_152, 8, 64, box=None, check=True)
    run(256, 10_000_000, 8, 64, check=True)
    run(512, 100_000_000, 8, 64, reps=3, check=False)
    run(512, 100_000_000, 8, 128, reps=2, check=False)
