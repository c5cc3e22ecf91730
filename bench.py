#!/usr/bin/env python
"""Benchmark of the hot path: batched periodic kNN queries against a kd-tree (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One "step" = one pass of the query path over one batch of synthetic queries: Morton ordering of the
batch (radix sort) + the kNN traversal kernel, results written to device memory.  Workload at every
N: BASELINE.json configs[2] -- 512^3 uniform points in the periodic unit box (tree built once on
rank 0 and replicated with one NCCL broadcast), k=8, 10^8 uniform random queries PER GPU (weak
scaling: queries are independent, every rank answers its own batch, no data-path collective).

Prints ONE JSON line (rank 0).  `value` = queries/s of the whole job with inputs resident in HBM;
`e2e` = the same metric through the host-pointer C-ABI call (nbk_tree_query) with pinned host
buffers, H2D/D2H inside the timed region; `roofline` = algorithmic bytes of the kNN kernel
(B_q = 12 + 8k + 16 V_n + 16 V_p) / its CUDA-event duration vs the measured HBM
peak; `cpu_baseline` = the reference's own CPU code (oracle/_ref) timed on this box's host cores on
a bounded sample.  `--impl reference` times only that CPU reference.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "kNN queries/sec (k=8, 512^3 periodic tree)"
UNIT = "queries/s"
# reference counters at leaf 64 for the headline config; re-measured live by the
# cpu_baseline leg and replaced when that leg runs
SURVEY_VP, SURVEY_VN = 255.8, 34.1


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--side", type=int, default=512, help="points = side^3 (512 = headline)")
    ap.add_argument("--queries", type=int, default=100_000_000, help="queries per GPU per step")
    ap.add_argument("-k", type=int, default=8)
    ap.add_argument("--leaf", type=int, default=64)
    ap.add_argument("--cpu-sample", type=int, default=2_000_000, help="queries of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def workload_name(args):
    return (f"{args.side}^3 uniform points, periodic unit box, leaf {args.leaf}, k={args.k}, "
            f"{args.queries:.0e} uniform random queries per GPU per step")


def algorithmic_bytes_per_query(k, v_n, v_p):
    return 12 + 8 * k + 16 * v_n + 16 * v_p


def ncu_traffic(args):
    """DRAM bytes (read + write) of one kNN-kernel launch from the committed `ncu --set full` capture
    of this workload (profiles/knn_traffic.json), or None when the run is not that workload."""
    path = os.path.join(ROOT, "profiles", "knn_traffic.json")
    if not os.path.exists(path):
        return None, "no ncu capture committed"
    with open(path) as f:
        rec = json.load(f)
    same = (rec.get("side") == args.side and rec.get("queries") == args.queries and rec.get("k") == args.k
            and rec.get("leaf") == args.leaf)
    if not same:
        return None, "ncu capture is for another workload"
    return rec["dram_bytes_per_launch"], rec.get("source", path)


def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---- clocks -------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi polled every 20 ms from before the warm-up; only the samples whose timestamp falls
    inside the timed region are reported (the recipe's clocks line)."""
    FIELDS = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []
        self.window = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def mark(self, t0: float, t1: float):
        """Wall-clock bounds of the timed region."""
        self.window = (t0, t1)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        self.thread.join(timeout=2)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = []
        for seen, line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                rows.append((seen, float(parts[1]), float(parts[2]), float(parts[3]), parts[4:8]))
            except ValueError:
                continue
        inside = [r for r in rows if self.window and self.window[0] <= r[0] <= self.window[1] + 0.03]
        used = inside or rows
        reasons = set()
        for r in used:
            for name, val in zip(names, r[4]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm = [r[1] for r in used]
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(r[2] for r in used) if used else None,
                "power_w_max": max(r[3] for r in used) if used else None,
                "samples": len(used), "samples_in_timed_region": len(inside), "reasons": sorted(reasons)}


# ---- CPU reference (the reference's own code, oracle/_ref; or the C restatement) -------------------
def cpu_reference_tree(points_host, leaf, box):
    from oracle import Oracle, Reference

    if Reference.available():
        return Reference.Tree(points_host, leaf, box), "reference", Reference.hardware_concurrency()
    return Oracle.Tree(points_host, leaf, box), "port", os.cpu_count() or 1


def run_reference_arm(args):
    """Times the reference's CPU kNN on this box: rank 0 only, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rng = np.random.Generator(np.random.Philox(42))
    n = args.side ** 3
    pts = rng.random((n, 3), dtype=np.float32)
    t0 = time.perf_counter()
    tree, kind, cores = cpu_reference_tree(pts, args.leaf, 1.0)
    build_s = time.perf_counter() - t0
    m = min(args.cpu_sample, args.queries)
    qrng = np.random.Generator(np.random.Philox(43))
    times = []
    for step in range(args.warmup + args.steps):
        q = qrng.random((m, 3), dtype=np.float32)
        t0 = time.perf_counter()
        tree.query(q, args.k, workers=0)
        dt = time.perf_counter() - t0
        if step >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = m * args.steps / total
    sample = f"{m} queries per step against the full {args.side}^3 tree, {cores} host threads (thread_pool chunks)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                         "build_seconds_1_thread": build_s, "build_mpts_per_s": n / build_s / 1e6},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(index: int):
    """Pins this process to the CPUs NVML names as local to GPU `index`, so that the pinned host buffers
    of the end-to-end leg are first-touched on that GPU's NUMA node (with 8 ranks copying 7.6 GB per
    step each, remote-node buffers halve the aggregate PCIe rate).  Best effort."""
    try:
        import pynvml

        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (os.cpu_count() + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        cpus = [c for c in cpus if c < os.cpu_count()]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return {"cpus": f"{cpus[0]}-{cpus[-1]}", "count": len(cpus)}
    except Exception as exc:  # pragma: no cover - depends on the box
        return {"error": str(exc)[:80]}
    return None


# ---- the B200 arm -----------------------------------------------------------------------------------
def run_b200_arm(args):
    import torch
    import torch.distributed as dist

    from nbodyhpc_b200 import capi
    from nbodyhpc_b200.dist import replicate_tree

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # (N = 1 keeps all cores: the CPU-baseline leg of that run uses every host thread)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 and not os.environ.get("NBK_BENCH_NO_BIND") else None
    if world > 1:
        # stdout carries exactly one JSON line: no NCCL version banner (NCCL_DEBUG=VERSION prints it there)
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    capi.lib()  # fail loudly if the extension is missing

    n, m, k = args.side ** 3, args.queries, args.k
    stream = torch.cuda.current_stream().cuda_stream

    # --- tree: built on rank 0, replicated with one NCCL broadcast --------------------------------
    tree, build_ms, build_all_ms, pts_host = None, None, None, None
    if rank == 0:
        g = torch.Generator(device=dev); g.manual_seed(42)
        pts = torch.rand((n, 3), device=dev, generator=g)
        # one untimed full-size build (first-touch of the scratch pool), then three timed ones
        capi.Tree.build_device(pts.data_ptr(), n, args.leaf, 1.0, stream=stream).close()
        build_all_ms = []
        for _ in range(3):
            if tree is not None:
                tree.close()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            tree = capi.Tree.build_device(pts.data_ptr(), n, args.leaf, 1.0, stream=stream)
            e1.record(); torch.cuda.synchronize()
            build_all_ms.append(e0.elapsed_time(e1))
        build_ms = float(np.median(build_all_ms))
        if world == 1 and not args.no_cpu_baseline:
            pts_host = pts.cpu().numpy()
        del pts
    bcast_ms = None
    if world > 1:
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        tree = replicate_tree(tree, src=0, device=local_rank)
        torch.cuda.synchronize(); dist.barrier()
        bcast_ms = 1e3 * (time.perf_counter() - t0)
    meta = tree.meta

    # --- this rank's query batch (resident in HBM) ---------------------------------------------------
    g = torch.Generator(device=dev); g.manual_seed(43 + rank)
    q = torch.rand((m, 3), device=dev, generator=g)
    out_d = torch.empty((m, k), device=dev, dtype=torch.float32)
    out_i = torch.empty((m, k), device=dev, dtype=torch.int32)

    def step():
        tree.query_device(q.data_ptr(), m, k, out_d.data_ptr(), out_i.data_ptr(), stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step()
    barrier()
    capi.profile_read(capi.SECTION_KNN_KERNEL); capi.profile_read(capi.SECTION_QUERY_ORDER)
    capi.profile_enable(True)
    launches0 = capi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    wall0 = time.time()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    sampler.mark(wall0, time.time())
    launches = capi.launch_count() - launches0
    capi.profile_enable(False)
    knn_ms, knn_cnt = capi.profile_read(capi.SECTION_KNN_KERNEL)
    order_ms, _ = capi.profile_read(capi.SECTION_QUERY_ORDER)
    clocks = sampler.stop() if rank == 0 else None
    elapsed = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
    total_ms = float(elapsed.item())
    value = world * m * args.steps / (total_ms * 1e-3)

    # --- end to end through the host-pointer C ABI (pinned host buffers, copies timed) -------------
    e2e = None
    if not args.no_e2e:
        q_host = torch.empty((m, 3), dtype=torch.float32, pin_memory=True)
        q_host.copy_(q)
        od_host = torch.empty((m, k), dtype=torch.float32, pin_memory=True)
        oi_host = torch.empty((m, k), dtype=torch.int32, pin_memory=True)
        e2e_steps = max(1, min(args.steps, 3))
        tree.query_raw(q_host.data_ptr(), m, k, od_host.data_ptr(), oi_host.data_ptr())  # warm-up
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            tree.query_raw(q_host.data_ptr(), m, k, od_host.data_ptr(), oi_host.data_ptr())
            checksum = float(od_host[0, 0])  # the result is read on the host
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * m * e2e_steps / float(dt.item()), "unit": UNIT, "steps": e2e_steps,
               "h2d_bytes_per_step": m * 12, "d2h_bytes_per_step": m * k * 8,
               "api": "nbk_tree_query (host pointers, pinned): slices of 2^21..2^24 queries on 3 streams, "
                      "H2D / kernel / D2H overlapped", "cpu_affinity": numa}
        same = bool(torch.equal(od_host[:100000], out_d[:100000].cpu()))
        e2e["matches_device_path"] = same
        del q_host, od_host, oi_host

    # --- CPU baseline: the reference's code on this box's cores, bounded sample (rank 0, N=1) -------
    cpu_baseline, v_n, v_p, parity = None, SURVEY_VN, SURVEY_VP, None
    counters_source = "SURVEY.md 8(d) (reference counters, 512^3 periodic k=8 leaf 64)"
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import compare_knn

        t0 = time.perf_counter()
        ref_tree, kind, cores = cpu_reference_tree(pts_host, args.leaf, 1.0)
        ref_build_s = time.perf_counter() - t0
        ms = min(args.cpu_sample, m)
        q_s = q[:ms].cpu().numpy()
        t0 = time.perf_counter()
        d_ref, i_ref, stats = ref_tree.query(q_s, k, workers=0, return_stats=True)
        ref_q_s = time.perf_counter() - t0
        v_n, v_p = float(stats[0]) / ms, float(stats[2]) / ms
        counters_source = f"reference KDTreeQueryStatistics measured on this run's {ms}-query sample"
        rep = compare_knn(out_d[:ms].cpu().numpy(), out_i[:ms].cpu().numpy().view(np.uint32), d_ref, i_ref,
                          pts_host, q_s, 1.0)
        parity = {"rows": rep.rows, "rows_equal": rep.rows_equal,
                  "rows_equal_after_tie_canonicalisation": rep.rows_equal_after_tie_canonicalisation,
                  "rows_boundary_tie_verified": rep.rows_boundary_tie_verified, "rows_wrong": rep.rows_wrong}
        cpu_baseline = {
            "value": ms / ref_q_s, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{ms} of the step's queries against the full {args.side}^3 reference tree, "
                      f"{cores} host threads (thread_pool chunks as pybind.cpp:164-172)",
            "build_seconds_1_thread": ref_build_s, "build_mpts_per_s": n / ref_build_s / 1e6,
            "nodes_visited_per_query": v_n, "points_visited_per_query": v_p,
        }

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_hbm_peak()
    traffic, traffic_src = ncu_traffic(args)
    b_q = algorithmic_bytes_per_query(k, v_n, v_p)
    # per-launch figures (one kNN launch pair per step today; written for any number)
    knn_ms_per_launch = knn_ms / max(knn_cnt, 1)
    queries_per_launch = m * args.steps / max(knn_cnt, 1)
    achieved = b_q * queries_per_launch / (knn_ms_per_launch * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "parallelism": f"tree replicated x{world} (NCCL broadcast), "
                   "queries sharded per GPU, no data-path collective",
                   "l2": "inputs larger than L2: 1.2 GB of queries + 2.2 GB tree per step, 6.4 GB written",
                   "tree": {"n_padded": int(meta.n_padded), "n_nodes": int(meta.n_nodes), "leaf_size": args.leaf}},
        "e2e": e2e,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {
            "bound": "hbm", "kernel": "knn_lane_kernel<K=8,periodic> (primary pass; the boundary pass is included in the timed section)", "achieved": achieved, "peak": peak,
            "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": peak_src,
            "algorithmic_bytes_per_query": b_q, "queries_per_launch": queries_per_launch,
            "launches_per_step": knn_cnt / args.steps, "counters": counters_source,
            "kernel_ms_per_launch": knn_ms_per_launch, "kernel_ms_per_step": knn_ms / args.steps,
            "kernel_share_of_step": knn_ms / total_ms,
            "query_order_ms_per_step": order_ms / args.steps,
        },
        "cpu_baseline": cpu_baseline,
        "parity_sample": parity,
        "build": {"ms": build_ms, "mpts_per_s": n / (build_ms * 1e-3) / 1e6 if build_ms else None,
                  "all_ms": build_all_ms, "timing": "median of 3 builds from device-resident points (CUDA events "
                  "around nbk_tree_build_device, host work included), after one untimed build",
                  "algorithmic_bytes": int(meta.n_levels) * int(meta.n_padded) * 32,
                  "roofline_frac": (int(meta.n_levels) * int(meta.n_padded) * 32 / (build_ms * 1e-3) / 1e9 / peak)
                  if build_ms else None,
                  "broadcast_ms": bcast_ms},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
