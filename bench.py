#!/usr/bin/env python
"""Benchmark of the hot path: batched periodic kNN queries against a kd-tree (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--scaling strong|weak] [--config 3|4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One "step" = one pass of the query path over one batch of synthetic queries: Morton ordering of the
batch (radix sort) + the kNN traversal kernel, results written to device memory.

Workload (default, --config 3): BASELINE.json configs[2] -- 512^3 uniform points in the periodic unit
box (tree built once on rank 0 and replicated with one NCCL broadcast), k=8, 10^8 uniform random
queries.  --scaling strong (default, what the north star asks for: "a k=8 query of 10^8 random points
... sharded over 1/2/4/8 GPUs"): the 10^8 queries are ONE batch, rank r answers its contiguous chunk
(thread_pool::parallelize_loop's split, thread_pool.hpp:163-179); --scaling weak: 10^8 queries PER GPU.
Queries are independent, so there is no data-path collective either way.
--config 4: BASELINE.json configs[3] -- 512^3 Zel'dovich-displaced lattice, periodic; a step is the fused
kNN-CDF (ks = 1,2,4,8,16,32: one traversal at k = 32, histograms accumulated on the device).

Prints ONE JSON line (rank 0).  `value` = queries/s of the whole job with inputs resident in HBM;
`e2e` = the same metric through the host-pointer C-ABI call (nbk_tree_query) with pinned host
buffers, H2D/D2H inside the timed region, and next to it the ceiling of this box's host<->device
copies measured in the same run (no kernels); `e2e_numpy` = the reference-facing Python call
KDTree.query(numpy) -> fresh numpy arrays (pageable both ways); `roofline` = algorithmic bytes of the
kNN kernel (B_q = 12 + 8k + 16 V_n + 16 V_p) / its CUDA-event duration vs the measured
HBM peak; `cpu_baseline` = the reference's own CPU code (oracle/_ref) timed on this box's host cores
on a bounded sample; `parity_sample` = the same sample compared row by row (squared distances bit for
bit, indices exactly) on rank 0 and on the last rank.  `--impl reference` times only the CPU reference.
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

UNIT = "queries/s"
CDF_KS = [1, 2, 4, 8, 16, 32]
# reference counters at leaf 64 for the headline config; re-measured live by the
# cpu_baseline leg and replaced when that leg runs
SURVEY_VP, SURVEY_VN = 255.8, 34.1


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=[3, 4], help="BASELINE.json configuration (1-based)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--side", type=int, default=512, help="points = side^3 (512 = headline)")
    ap.add_argument("--queries", type=int, default=100_000_000,
                    help="queries per step: of the whole job (strong) or per GPU (weak)")
    ap.add_argument("-k", type=int, default=None, help="default 8 (config 3) / 32 (config 4: the CDF's largest k)")
    ap.add_argument("--leaf", type=int, default=64)
    ap.add_argument("--cpu-sample", type=int, default=2_000_000, help="queries of the CPU baseline / parity sample")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU reference legs (baseline + parity)")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.k is None:
        args.k = 8 if args.config == 3 else max(CDF_KS)
    if args.config == 4 and args.k != max(CDF_KS):
        ap.error(f"--config 4 runs the fused CDF at k = {max(CDF_KS)}")
    return args


def metric_name(args):
    if args.config == 4:
        return f"kNN-CDF queries/sec (k=1..{max(CDF_KS)}, {args.side}^3 clustered periodic tree)"
    return f"kNN queries/sec (k={args.k}, {args.side}^3 periodic tree)"


def workload_name(args, world=1):
    pts = "uniform" if args.config == 3 else "Zel'dovich-displaced lattice (rms 1.5 cells, P(k) ~ k^-2)"
    per = "of the whole job, sharded in contiguous chunks" if args.scaling == "strong" else "per GPU"
    what = f"k={args.k}" if args.config == 3 else f"fused kNN-CDF ks={CDF_KS} (one traversal at k={max(CDF_KS)})"
    return (f"{args.side}^3 {pts} points, periodic unit box, leaf {args.leaf}, {what}, "
            f"{args.queries:.0e} uniform random queries per step {per}")


def algorithmic_bytes_per_query(k, v_n, v_p, rows=True):
    return 12 + (8 * k if rows else 0) + 16 * v_n + 16 * v_p


def ncu_traffic(args):
    """DRAM bytes (read + write) of one kNN-kernel launch from the committed `ncu --set full` capture
    of this workload (profiles/knn_traffic.json), or None when the run is not that workload."""
    path = os.path.join(ROOT, "profiles", "knn_traffic.json")
    if not os.path.exists(path):
        return None, "no ncu capture committed"
    with open(path) as f:
        rec = json.load(f)
    same = (args.config == 3 and rec.get("side") == args.side and rec.get("queries") == args.queries
            and rec.get("k") == args.k and rec.get("leaf") == args.leaf)
    if not same:
        return None, "ncu capture is for another workload"
    # the capture names the kernel source it was taken from: say so when the kernels have changed since
    import hashlib

    with open(os.path.join(ROOT, "nbodyhpc_b200", "csrc", "knn_query.cuh"), "rb") as f:
        current = hashlib.sha1(f.read()).hexdigest()
    note = rec.get("source", path)
    if rec.get("kernel_source_sha1") != current:
        note = "STALE (knn_query.cuh changed since the capture; rerun scripts/refresh_traffic.py): " + note
    return rec["dram_bytes_per_launch"], note


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


def host_points(args):
    """The benchmark's particle set on the host without a GPU (reference arm): uniform from numpy's Philox;
    the clustered set needs FFTs of a 512^3 field and is generated with torch on the CPU."""
    n = args.side ** 3
    if args.config == 3:
        return np.random.Generator(np.random.Philox(42)).random((n, 3), dtype=np.float32)
    import torch

    from scripts.synthetic import zeldovich

    return zeldovich(args.side, 42, torch.device("cpu")).numpy()


def run_reference_arm(args):
    """Times the reference's CPU kNN on this box: rank 0 only, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.side ** 3
    pts = host_points(args)
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
    sample = (f"{m} queries per step against the full {args.side}^3 tree, {cores} host threads (thread_pool chunks)"
              + ("" if args.config == 3 else f"; rows at k={args.k}, what the kNN-CDF consumes"))
    line = {
        "impl": "reference", "metric": metric_name(args), "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
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


def host_copy_ceiling(torch, dev, h2d_bytes, d2h_bytes, reps, reduce_max):
    """What this box's host<->device copies allow with NO kernels: one step's H2D and D2H bytes moved
    concurrently between pinned host memory and the device in 64 MB slices, one cudaMemcpyAsync per slice,
    one stream per direction; every rank at once, max over ranks.  Returns seconds per step."""
    slice_b = 64 << 20
    pin_in = torch.empty(max(h2d_bytes, 1), dtype=torch.uint8, pin_memory=True)
    pin_out = torch.empty(max(d2h_bytes, 1), dtype=torch.uint8, pin_memory=True)
    d_in = torch.empty(max(h2d_bytes, 1), dtype=torch.uint8, device=dev)
    d_out = torch.empty(max(d2h_bytes, 1), dtype=torch.uint8, device=dev)
    pin_in.zero_(); pin_out.zero_()
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def once():
        with torch.cuda.stream(s_in):
            for b in range(0, h2d_bytes, slice_b):
                d_in[b:b + slice_b].copy_(pin_in[b:b + slice_b], non_blocking=True)
        with torch.cuda.stream(s_out):
            for b in range(0, d2h_bytes, slice_b):
                pin_out[b:b + slice_b].copy_(d_out[b:b + slice_b], non_blocking=True)

    once(); torch.cuda.synchronize()
    reduce_max(0.0)  # line the ranks up
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    torch.cuda.synchronize()
    return reduce_max((time.perf_counter() - t0) / reps)


# ---- the B200 arm -----------------------------------------------------------------------------------
def run_b200_arm(args):
    import torch
    import torch.distributed as dist

    from nbodyhpc_b200 import capi
    from nbodyhpc_b200.dist import arena_tensor, replicate_tree, shard_range
    from scripts.synthetic import uniform, zeldovich

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # (N = 1 keeps all cores: the CPU-baseline leg of that run uses every host thread)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 and not os.environ.get("NBK_BENCH_NO_BIND") else None
    comm_setup_ms = None
    if world > 1:
        # stdout carries exactly one JSON line: no NCCL version banner (NCCL_DEBUG=VERSION prints it there)
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
        # the first collective creates the communicator: time that apart from the tree broadcast
        t0 = time.perf_counter()
        dist.broadcast(torch.zeros(1, device=dev), src=0)
        torch.cuda.synchronize()
        comm_setup_ms = 1e3 * (time.perf_counter() - t0)
    capi.lib()  # fail loudly if the extension is missing

    def reduce_max(x: float) -> float:
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n, k = args.side ** 3, args.k
    if args.scaling == "strong":
        q_begin, q_end = shard_range(args.queries, rank, world)
        m_job = args.queries
    else:
        q_begin, q_end = rank * args.queries, (rank + 1) * args.queries
        m_job = args.queries * world
    m = q_end - q_begin
    stream = torch.cuda.current_stream().cuda_stream
    last = world - 1
    want_checks = not args.no_cpu_baseline
    checker_ranks = {0, last} if want_checks else set()

    def make_points():
        return uniform(n, 42, dev) if args.config == 3 else zeldovich(args.side, 42, dev)

    # --- tree: built on rank 0, replicated with one NCCL broadcast --------------------------------
    tree, build_ms, build_all_ms, first_ms, pts_host = None, None, None, None, None
    if rank == 0:
        pts = make_points()
        torch.cuda.synchronize()
        # the first build of the process (cold: scratch blocks, function attributes), then three timed ones
        t0 = time.perf_counter()
        capi.Tree.build_device(pts.data_ptr(), n, args.leaf, 1.0, stream=stream).close()
        torch.cuda.synchronize()
        first_ms = 1e3 * (time.perf_counter() - t0)
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
        if rank in checker_ranks:
            pts_host = pts.cpu().numpy()
        del pts
    elif rank in checker_ranks:
        pts_host = make_points().cpu().numpy()  # same generator, same seed: rank 0's points
    bcast_ms, bcast_repeat_ms, replicas_identical = None, None, None
    if world > 1:
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        tree = replicate_tree(tree, src=0, device=local_rank)
        torch.cuda.synchronize(); dist.barrier()
        bcast_ms = 1e3 * (time.perf_counter() - t0)
        # the same bytes once more: NCCL's channels and buffers for large messages are set up by now
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        dist.broadcast(arena_tensor(tree), src=0)
        torch.cuda.synchronize(); dist.barrier()
        bcast_repeat_ms = 1e3 * (time.perf_counter() - t0)
        # byte-identical replicas: a checksum of every rank's arena against rank 0's
        words = arena_tensor(tree).view(torch.int64)
        digest = torch.stack([words.sum(), (words * torch.arange(1, words.numel() + 1, device=dev)).sum()])
        ref_digest = digest.clone()
        dist.broadcast(ref_digest, src=0)
        same = torch.tensor([int(torch.equal(digest, ref_digest))], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        replicas_identical = bool(same.item())
        del words
    meta = tree.meta

    # --- this rank's chunk of the query batch (resident in HBM) -------------------------------------
    # one generator stream for the whole job, so that the batch does not depend on the number of ranks:
    # chunks of 2^24 queries, chunk c drawn with seed 43 + c; a rank draws the chunks its range touches
    chunk = 1 << 24
    q = torch.empty((m, 3), device=dev)
    for c in range(q_begin // chunk, (max(q_end, 1) - 1) // chunk + 1):
        block = uniform(chunk, 43 + c, dev)
        lo, hi = max(q_begin, c * chunk), min(q_end, (c + 1) * chunk)
        if hi > lo:
            q[lo - q_begin:hi - q_begin] = block[lo - c * chunk:hi - c * chunk]
        del block
    rows = args.config == 3
    if rows:
        out_d = torch.empty((m, k), device=dev, dtype=torch.float32)
        out_i = torch.empty((m, k), device=dev, dtype=torch.int32)
    else:
        edges = np.concatenate([[0.0], np.geomspace(2e-4, 0.05, 48)]).astype(np.float32)
        d_edges = torch.from_numpy(edges).to(dev)
        d_counts = torch.zeros((len(CDF_KS), len(edges) - 1), device=dev, dtype=torch.int64)

    def step():
        if rows:
            tree.query_device(q.data_ptr(), m, k, out_d.data_ptr(), out_i.data_ptr(), stream)
        else:
            tree.knn_cdf_device(q.data_ptr(), m, CDF_KS, d_edges.data_ptr(), len(edges) - 1, d_counts.data_ptr(), stream)

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
    total_ms = reduce_max(e0.elapsed_time(e1))
    value = m_job * args.steps / (total_ms * 1e-3)
    knn_ms_max = reduce_max(knn_ms)

    # --- config 4: the row queries the fused CDF replaces, one timed pass per k ----------------------
    per_k = None
    if not rows:
        per_k = {}
        for kk in CDF_KS:
            od = torch.empty((m, kk), device=dev, dtype=torch.float32)
            oi = torch.empty((m, kk), device=dev, dtype=torch.int32)
            tree.query_device(q.data_ptr(), m, kk, od.data_ptr(), oi.data_ptr(), stream)
            barrier()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            tree.query_device(q.data_ptr(), m, kk, od.data_ptr(), oi.data_ptr(), stream)
            a1.record(); barrier()
            per_k[str(kk)] = m_job / (reduce_max(a0.elapsed_time(a1)) * 1e-3)
            if kk == k:
                out_d, out_i = od, oi  # the parity leg compares these rows
            else:
                del od, oi

    # --- end to end through the host-pointer C ABI (pinned host buffers, copies timed) -------------
    e2e, e2e_numpy = None, None
    if not args.no_e2e and rows:
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
        dt = reduce_max(time.perf_counter() - t0)
        e2e = {"value": m_job * e2e_steps / dt, "unit": UNIT, "steps": e2e_steps,
               "h2d_bytes_per_step": m * 12, "d2h_bytes_per_step": m * k * 8,
               "api": "nbk_tree_query (host pointers, pinned): slices of 2^19, 2^20, 2^21, then 2^22 queries on 3 "
                      "streams, H2D / kernel / D2H overlapped", "cpu_affinity": numa}
        same = bool(torch.equal(od_host[:100000], out_d[:100000].cpu()))
        e2e["matches_device_path"] = same
        del q_host, od_host, oi_host
        # the box's own ceiling for exactly these copies, no kernels (every rank at once)
        ceiling_s = host_copy_ceiling(torch, dev, m * 12, m * k * 8, 2, reduce_max)
        e2e["host_ceiling"] = {"value": m_job / ceiling_s, "unit": UNIT,
                               "gbs_d2h_aggregate": world * m * k * 8 / ceiling_s / 1e9,
                               "how": "this step's H2D + D2H bytes between pinned host memory and the device, both "
                                      "directions at once, 64 MB slices (one cudaMemcpyAsync each), no kernels, all "
                                      "ranks at once, max over ranks"}
        e2e["frac_of_host_ceiling"] = e2e["value"] / e2e["host_ceiling"]["value"]
        # the call a user of the reference API makes: numpy in (pageable), fresh numpy arrays out
        if world == 1:
            from nbodyhpc.kdtree import KDTree

            q_np = q.cpu().numpy()
            py_tree = KDTree(make_points(), leafsize=args.leaf, boxsize=1.0)
            t0 = time.perf_counter()
            d_np, i_np = py_tree.query(q_np, k=k)  # the first call: fresh pageable result arrays, staged copies
            first_s = time.perf_counter() - t0
            same_np = bool(np.array_equal(d_np[:100000], out_d[:100000].cpu().numpy()))
            del d_np, i_np  # a caller that drops its results: the buffers are recycled, page-locked in the background
            time.sleep(1.5)
            times = []
            for _ in range(3):
                t0 = time.perf_counter()
                d_np, i_np = py_tree.query(q_np, k=k)
                checksum = float(d_np[0, 0])
                times.append(time.perf_counter() - t0)
                del d_np, i_np
            e2e_numpy = {"value": m / float(np.median(times)), "unit": UNIT, "seconds": times,
                         "first_call": {"value": m / first_s, "seconds": first_s},
                         "matches_device_path": same_np,
                         "api": "nbodyhpc.kdtree.KDTree.query(numpy (M,3) float32, k) -> new numpy (M,k) float32 + uint32 "
                                "arrays; queries pageable (staged through the library's pinned ring); value = median of 3 "
                                "calls whose result buffers are recycled from results the caller dropped (page-locked by a "
                                "background thread, written by the copy engine directly); first_call = fresh pageable "
                                "result arrays",
                         "host_path": capi.host_path_stats()}
            del py_tree, q_np

    # --- CPU legs: the reference's code on this box's cores, bounded sample ---------------------------
    # rank 0: baseline timing + parity; last rank: parity of ITS chunk on ITS replica
    cpu_baseline, v_n, v_p, parity = None, SURVEY_VN, SURVEY_VP, {}
    counters_source = "SURVEY.md 8(d) (reference counters, 512^3 periodic k=8 leaf 64)"
    if rank in checker_ranks:
        from oracle import compare_knn

        t0 = time.perf_counter()
        ref_tree, kind, cores = cpu_reference_tree(pts_host, args.leaf, 1.0)
        ref_build_s = time.perf_counter() - t0
        ms = min(args.cpu_sample, m)
        q_s = q[:ms].cpu().numpy()
        t0 = time.perf_counter()
        d_ref, i_ref, stats = ref_tree.query(q_s, k, workers=0, return_stats=True)
        ref_q_s = time.perf_counter() - t0
        rep = compare_knn(out_d[:ms].cpu().numpy(), out_i[:ms].cpu().numpy().view(np.uint32), d_ref, i_ref,
                          pts_host, q_s, 1.0)
        # and the squared distances, bit for bit (what the search ranks by; sqrt is many-to-one)
        sq_d = torch.empty((ms, k), device=dev, dtype=torch.float32)
        sq_i = torch.empty((ms, k), device=dev, dtype=torch.int32)
        tree.query_device(q.data_ptr(), ms, k, sq_d.data_ptr(), sq_i.data_ptr(), stream, squared=True)
        torch.cuda.synchronize()
        d2_ref, i2_ref = ref_tree.query(q_s, k, workers=0, squared=True)
        rep2 = compare_knn(sq_d.cpu().numpy(), sq_i.cpu().numpy().view(np.uint32), d2_ref, i2_ref, pts_host, q_s, 1.0,
                           squared=True)
        mine = {"rank": rank, "rows": rep.rows, "rows_equal": rep.rows_equal,
                "rows_equal_after_tie_canonicalisation": rep.rows_equal_after_tie_canonicalisation,
                "rows_boundary_tie_verified": rep.rows_boundary_tie_verified, "rows_wrong": rep.rows_wrong,
                "squared_rows_equal": rep2.rows_equal, "squared_rows_wrong": rep2.rows_wrong,
                "queries": f"[{q_begin}, {q_begin + ms}) of the job's batch"}
        if not rows:
            # the fused histogram of this sample == numpy.histogram of the reference's rows
            h = tree.knn_cdf(q_s, CDF_KS, edges)
            mine["cdf_equals_histogram_of_reference_rows"] = all(
                np.array_equal(h[r], np.histogram(d_ref[:, kk - 1], edges)[0].astype(np.uint64))
                for r, kk in enumerate(CDF_KS))
        parity[rank] = mine
        if rank == 0:
            v_n, v_p = float(stats[0]) / ms, float(stats[2]) / ms
            counters_source = f"reference KDTreeQueryStatistics measured on this run's {ms}-query sample"
            if world == 1:
                cpu_baseline = {
                    "value": ms / ref_q_s, "unit": UNIT, "cores": cores, "kind": kind,
                    "sample": f"{ms} of the step's queries against the full {args.side}^3 reference tree, "
                              f"{cores} host threads (thread_pool chunks as pybind.cpp:164-172)",
                    "build_seconds_1_thread": ref_build_s, "build_mpts_per_s": n / ref_build_s / 1e6,
                    "nodes_visited_per_query": v_n, "points_visited_per_query": v_p,
                }
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, parity)
        parity = {r: p for g in gathered for r, p in g.items()}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_hbm_peak()
    traffic, traffic_src = ncu_traffic(args)
    b_q = algorithmic_bytes_per_query(k, v_n, v_p, rows)
    # per-launch figures of rank 0's kernel (the slowest rank's time is reported next to it)
    knn_ms_per_launch = knn_ms / max(knn_cnt, 1)
    queries_per_launch = m * args.steps / max(knn_cnt, 1)
    achieved = b_q * queries_per_launch / (knn_ms_per_launch * 1e-3) / 1e9
    per = "of the whole job, contiguous chunk per rank" if args.scaling == "strong" else "per GPU"
    line = {
        "metric": metric_name(args), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args, world), "baseline_config": args.config,
                   "queries_per_gpu_per_step": m, "queries_per_step": m_job,
                   "parallelism": f"tree replicated x{world} (NCCL broadcast), {args.queries:.0e} queries {per}, "
                                  "no data-path collective",
                   "l2": f"inputs larger than L2: {m * 12 / 1e9:.2f} GB of queries + {meta.arena_bytes / 1e9:.1f} GB tree "
                         f"per GPU per step" + (f", {m * k * 8 / 1e9:.2f} GB written" if rows else ""),
                   "tree": {"n_padded": int(meta.n_padded), "n_nodes": int(meta.n_nodes), "leaf_size": args.leaf}},
        "e2e": e2e,
        "e2e_numpy": e2e_numpy,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {
            "bound": "hbm", "kernel": f"knn_lane_kernel<K={k},periodic> (primary pass; the boundary pass is included in "
                                      "the timed section)", "achieved": achieved, "peak": peak,
            "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": peak_src,
            "algorithmic_bytes_per_query": b_q, "queries_per_launch": queries_per_launch,
            "launches_per_step": knn_cnt / args.steps, "counters": counters_source,
            "kernel_ms_per_launch": knn_ms_per_launch, "kernel_ms_per_step": knn_ms / args.steps,
            "kernel_ms_per_step_slowest_rank": knn_ms_max / args.steps,
            "kernel_share_of_step": knn_ms / total_ms,
            "query_order_ms_per_step": order_ms / args.steps,
        },
        "cpu_baseline": cpu_baseline,
        "parity_sample": [parity[r] for r in sorted(parity)] if parity else None,
        "rows_per_k": per_k,
        "build": {"ms": build_ms, "mpts_per_s": n / (build_ms * 1e-3) / 1e6 if build_ms else None,
                  "all_ms": build_all_ms, "first_ms": first_ms,
                  "timing": "median of 3 builds from device-resident points (CUDA events around nbk_tree_build_device, "
                            "host work included); first_ms = the first build of the process, host clock",
                  "algorithmic_bytes": int(meta.n_levels) * int(meta.n_padded) * 32,
                  "roofline_frac": (int(meta.n_levels) * int(meta.n_padded) * 32 / (build_ms * 1e-3) / 1e9 / peak)
                  if build_ms else None,
                  "broadcast_ms": bcast_ms, "comm_setup_ms": comm_setup_ms, "broadcast_repeat_ms": bcast_repeat_ms,
                  "broadcast_repeat_gbs": meta.arena_bytes / (bcast_repeat_ms * 1e-3) / 1e9 if bcast_repeat_ms else None,
                  "replicas_identical": replicas_identical},
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
