#!/usr/bin/env python
"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
    python scripts/launch_summary.py launches.csv [first-kernel-substring]"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    start_at = sys.argv[2] if len(sys.argv) > 2 else None
    with open(path) as f:
        lines = [line for line in f if not line.startswith("==")]
    seq = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = row["Kernel Name"].split("(")[0].replace("void ", "").replace("nbk::", "")
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v *= {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}[unit]
        seq.append((name, v))
    if start_at:
        idx = [i for i, (n, _) in enumerate(seq) if start_at in n]
        seq = seq[idx[-1]:] if idx else seq
    agg = collections.OrderedDict()
    for n, v in seq:
        a = agg.setdefault(n, [0, 0.0, []])
        a[0] += 1
        a[1] += v
        a[2].append(v)
    total = sum(v for _, v in seq)
    print(f"{'kernel':64s} {'launches':>8s} {'total ms':>10s} {'share':>7s}   per-launch ms (first..last)")
    for n, (c, v, each) in agg.items():
        shown = " ".join(f"{e:.3f}" for e in (each if len(each) <= 16 else each[:8] + each[-8:]))
        print(f"{n:64s} {c:8d} {v:10.3f} {100 * v / total:6.1f}%   {shown}")
    print(f"{'total':64s} {len(seq):8d} {total:10.3f}")


if __name__ == "__main__":
    main()
