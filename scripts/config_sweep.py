"""Runs the BASELINE.json configurations on one B200 and records build rate, query throughput and
full-size self-consistency checks (sorted rows, distances bit-equal to a torch recomputation of the
returned indices, exhaustive check of a query sample on the device).

    python scripts/config_sweep.py [--configs 1,2,3,4,5] [--out gpurun_out/config_sweep.json]

torch is only the harness here (synthetic data, FFTs for the Zel'dovich field, checks); every
product call goes through the C ABI (nbodyhpc_b200.capi).
"""
import argparse
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from nbodyhpc_b200 import capi  # noqa: E402

DEV = torch.device("cuda", 0)


from scripts.synthetic import uniform as _uniform, zeldovich as _zeldovich  # noqa: E402


def uniform(n, seed):
    return _uniform(n, seed, DEV)


def zeldovich(side, seed, rms_cells=1.5, index=-2.0):
    return _zeldovich(side, seed, DEV, rms_cells, index)


def d2_torch(p, q, box):
    """Reference arithmetic with separate (unfused) torch ops."""
    acc = None
    for a in range(3):
        d = p[..., a] - q[..., a]
        t = d * d
        if box is not None:
            dp, dm = d + box, d - box
            t = torch.minimum(torch.minimum(t, dp * dp), dm * dm)
        acc = t if acc is None else acc + t
    return acc


def check(points, q, out_d, out_i, k, box, n_real, sample=256, chunk=None):
    m = q.shape[0]
    chunk = chunk or max(1, (1 << 25) // k)
    res = {"rows": m}
    bad_sorted = bad_dist = 0
    for b in range(0, m, chunk):
        e = min(m, b + chunk)
        d = out_d[b:e]; i = out_i[b:e].long()
        bad_sorted += int((d[:, 1:] < d[:, :-1]).any(dim=1).sum()) if k > 1 else 0
        valid = (i >= 0) & (i < n_real)
        pi = points[i.clamp(min=0, max=n_real - 1)]
        dd = torch.sqrt(d2_torch(pi, q[b:e, None, :], box))
        bad_dist += int(((dd != d) & valid).any(dim=1).sum())
    res["rows_not_sorted"] = bad_sorted
    res["rows_distance_mismatch"] = bad_dist
    # exhaustive check of a sample: the k smallest (d2, index) over ALL points
    g = torch.Generator(device=DEV); g.manual_seed(7)
    pick = torch.randint(0, m, (sample,), device=DEV, generator=g)
    wrong = 0
    for j in pick.tolist():
        d2 = d2_torch(points, q[j][None, :], box)
        key = (d2.view(torch.int32).long() << 32) | torch.arange(points.shape[0], device=DEV)
        best = torch.topk(key, k, largest=False).values
        got = (out_d[j] ** 2)  # not used for equality (sqrt is not invertible); compare via indices + distances
        exp_i = (best & 0xFFFFFFFF)
        exp_d = torch.sqrt((best >> 32).int().view(torch.float32))
        if not (torch.equal(exp_i, out_i[j].long() & 0xFFFFFFFF) and torch.equal(exp_d, out_d[j])):
            wrong += 1
    res["exhaustive_sample"] = sample
    res["exhaustive_sample_wrong"] = wrong
    return res


def run_cdf(tree, queries, ks, rows_kmax, kmax, stream):
    """Fused kNN-CDF (nbk_tree_knn_cdf_device) vs the histogram of the rows of the k = max(ks) query."""
    m = queries.shape[0]
    hi = float(rows_kmax[:, kmax - 1].max())
    edges = torch.cat([torch.zeros(1, device=DEV), torch.logspace(np.log10(hi) - 3.0, np.log10(hi), 64, device=DEV)])
    edges[-1] = hi
    edges = edges.float().contiguous()
    n_bins = edges.numel() - 1
    times = []
    for _ in range(3):
        counts = torch.zeros((len(ks), n_bins), device=DEV, dtype=torch.int64)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        tree.knn_cdf_device(queries.data_ptr(), m, ks, edges.data_ptr(), n_bins, counts.data_ptr(), stream)
        b.record(); torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    ok = True
    for row, k in enumerate(ks):
        d = rows_kmax[:, k - 1].contiguous()
        # numpy.histogram semantics: [e_b, e_b+1), last bin closed
        b_idx = torch.searchsorted(edges, d, right=True) - 1
        b_idx[d == edges[-1]] = n_bins - 1
        inside = (b_idx >= 0) & (b_idx < n_bins) & (d <= edges[-1])
        expect = torch.bincount(b_idx[inside], minlength=n_bins)
        ok = ok and bool(torch.equal(expect, counts[row]))
    return {"cdf_ks": ks, "bins": n_bins, "ms": min(times), "mq_per_s": m / min(times) / 1e3,
            "equals_histogram_of_rows": ok}


def run(name, points, queries, ks, leaf, box, do_check=True, reps=3, cdf=False):
    n, m = points.shape[0], queries.shape[0]
    stream = torch.cuda.current_stream().cuda_stream
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    tree = capi.Tree.build_device(points.data_ptr(), n, leaf, box, stream=stream)
    e1.record(); torch.cuda.synchronize()
    build_ms = e0.elapsed_time(e1)
    meta = tree.meta
    out = {"config": name, "n_points": n, "n_queries": m, "leaf": leaf, "box": box, "build_ms": build_ms,
           "build_mpts_per_s": n / build_ms / 1e3, "n_nodes": int(meta.n_nodes), "n_levels": int(meta.n_levels),
           "queries": []}
    for k in ks:
        od = torch.empty((m, k), device=DEV, dtype=torch.float32)
        oi = torch.empty((m, k), device=DEV, dtype=torch.int32)
        times = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            tree.query_device(queries.data_ptr(), m, k, od.data_ptr(), oi.data_ptr(), stream)
            b.record(); torch.cuda.synchronize()
            times.append(a.elapsed_time(b))
        row = {"k": k, "ms": min(times), "mq_per_s": m / min(times) / 1e3}
        if do_check:
            row["check"] = check(points, queries, od, oi, k, box, n)
        out["queries"].append(row)
        print(json.dumps({"config": name, **row}), flush=True)
        if cdf and k == max(ks):
            out["cdf"] = run_cdf(tree, queries, ks, od, k, stream)
            print(json.dumps({"config": name, **out["cdf"]}), flush=True)
        del od, oi
    tree.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,2,3,4")
    ap.add_argument("--out", default="gpurun_out/config_sweep.json")
    args = ap.parse_args()
    todo = set(args.configs.split(","))
    results = []
    if "1" in todo:  # 1M uniform periodic, k=8, 1M queries
        results.append(run("1: 1M uniform periodic, 1M queries", uniform(1_000_000, 42), uniform(1_000_000, 43), [8], 64, 1.0))
    if "2" in todo:  # 128^3 non-periodic self-query k=1..16
        p = uniform(128 ** 3, 42)
        results.append(run("2: 128^3 open self-query", p, p, [1, 2, 4, 8, 16], 64, None))
    if "3" in todo:  # headline
        results.append(run("3: 512^3 uniform periodic, 1e8 queries", uniform(512 ** 3, 42), uniform(100_000_000, 43), [8], 64, 1.0))
    if "4" in todo:  # clustered
        p = zeldovich(512, 42)
        results.append(run("4: 512^3 Zel'dovich periodic, 1e8 uniform queries", p, uniform(100_000_000, 43),
                           [1, 2, 4, 8, 16, 32], 64, 1.0, cdf=True))
        del p
    if "5" in todo:  # HBM sizing
        p = uniform(1024 ** 3, 42)
        results.append(run("5: 1024^3 uniform periodic, 1e8 queries k=64", p, uniform(100_000_000, 43), [64], 64, 1.0,
                           reps=2))
    with open(args.out, "w") as f:
        json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
