"""Host<->device copy ceiling of this box for the end-to-end leg of bench.py, no kernels involved.

    python scripts/pcie_probe.py [--out profiles/r2_pcie_ceiling.json] [--gb 2]

For N = 1, 2, 4, 8 GPUs (those present): pinned host buffers, one stream per GPU and direction, 64 MB
slices, ONE cudaMemcpyAsync per slice (torch's non_blocking copy_ between a pinned and a device tensor),
all GPUs driven at once.  Three patterns: device->host only, host->device only, and both at once with the
byte ratio of the k=8 query (64 B of results out per 12 B of query in).  Also one GPU with PAGEABLE memory,
the case the staging ring of host_stage.cuh exists for.  Rates are aggregate GB/s over all GPUs.
"""
import argparse
import json
import os
import time

import torch


def run(n_gpus, d2h_bytes, h2d_bytes, pinned=True, reps=3, slice_b=64 << 20):
    devs = [torch.device("cuda", i) for i in range(n_gpus)]
    bufs = []
    for d in devs:
        with torch.cuda.device(d):
            bufs.append(dict(
                d_out=torch.empty(max(d2h_bytes, 1), dtype=torch.uint8, device=d),
                d_in=torch.empty(max(h2d_bytes, 1), dtype=torch.uint8, device=d),
                h_out=torch.empty(max(d2h_bytes, 1), dtype=torch.uint8, pin_memory=pinned).zero_(),
                h_in=torch.empty(max(h2d_bytes, 1), dtype=torch.uint8, pin_memory=pinned).zero_(),
                s_out=torch.cuda.Stream(d), s_in=torch.cuda.Stream(d)))

    def once():
        # interleave the GPUs slice by slice so that every link has work from the start
        for b in range(0, max(d2h_bytes, h2d_bytes), slice_b):
            for buf in bufs:
                if b < d2h_bytes:
                    with torch.cuda.stream(buf["s_out"]):
                        buf["h_out"][b:b + slice_b].copy_(buf["d_out"][b:b + slice_b], non_blocking=True)
                if b < h2d_bytes:
                    with torch.cuda.stream(buf["s_in"]):
                        buf["d_in"][b:b + slice_b].copy_(buf["h_in"][b:b + slice_b], non_blocking=True)

    def sync():
        for d in devs:
            torch.cuda.synchronize(d)

    once(); sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    sync()
    dt = (time.perf_counter() - t0) / reps
    return {"seconds": dt, "d2h_gbs": n_gpus * d2h_bytes / dt / 1e9, "h2d_gbs": n_gpus * h2d_bytes / dt / 1e9}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/pcie_ceiling.json")
    ap.add_argument("--gb", type=float, default=2.0, help="GB moved device->host per GPU and repetition")
    args = ap.parse_args()
    n_all = torch.cuda.device_count()
    big = int(args.gb * 1e9) // (64 << 20) * (64 << 20)
    small = big * 12 // 64 // 4096 * 4096
    res = {"gpu": torch.cuda.get_device_name(0), "gpus_present": n_all, "host_cpus": os.cpu_count(),
           "bytes_d2h_per_gpu": big, "slice_bytes": 64 << 20,
           "method": "pinned host memory, one stream per GPU and direction, one cudaMemcpyAsync per 64 MB slice, "
                     "all GPUs at once; aggregate GB/s", "by_gpus": {}}
    for n in (1, 2, 4, 8):
        if n > n_all:
            break
        res["by_gpus"][str(n)] = {
            "d2h_only": run(n, big, 0),
            "h2d_only": run(n, 0, big),
            "query_mix_64_to_12": run(n, big, small),
        }
        print(n, json.dumps(res["by_gpus"][str(n)]), flush=True)
    res["pageable_1gpu"] = {"d2h_only": run(1, big, 0, pinned=False, reps=1), "h2d_only": run(1, 0, big, pinned=False, reps=1)}
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
