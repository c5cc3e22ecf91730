"""Writes profiles/knn_traffic.json from an `ncu --set full` capture of the bench command's kNN kernel:

    ncu --set full --clock-control none --import-source on -k regex:knn_lane_kernel -s 2 -c 1 -f -o gpurun_out/knn \\
        python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e
    python scripts/refresh_traffic.py gpurun_out/knn.ncu-rep

The record carries the SHA-1 of the query kernels' source (nbodyhpc_b200/csrc/knn_query.cuh) at capture time;
bench.py reports `traffic_stale` when the source has changed since."""
import csv
import hashlib
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def kernel_source_sha1():
    with open(os.path.join(ROOT, "nbodyhpc_b200", "csrc", "knn_query.cuh"), "rb") as f:
        return hashlib.sha1(f.read()).hexdigest()


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, row = rows[0], rows[1], rows[2]

    def metric(name):
        i = hdr.index(name)
        v, u = float(row[i].replace(",", "")), units[i]
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "msecond": 1.0,
                 "usecond": 1e-3, "second": 1e3, "ns": 1e-6, "nsecond": 1e-6}.get(u, 1.0)
        return v * scale

    rd, wr = metric("dram__bytes_read.sum"), metric("dram__bytes_write.sum")
    rec = {
        "side": 512, "queries": 100000000, "k": 8, "leaf": 64,
        "kernel": row[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "knn_lane_kernel",
        "dram_bytes_read": int(rd), "dram_bytes_write": int(wr), "dram_bytes_per_launch": int(rd + wr),
        "gpu_time_ms_under_ncu": metric("gpu__time_duration.sum"),
        "kernel_source_sha1": kernel_source_sha1(),
        "source": "ncu --set full --clock-control none --import-source on -k regex:knn_lane_kernel -s 2 -c 1 python bench.py "
                  "--steps 2 --warmup 1 --no-cpu-baseline --no-e2e: dram__bytes_read.sum + dram__bytes_write.sum "
                  "(scripts/refresh_traffic.py)",
    }
    with open(os.path.join(ROOT, "profiles", "knn_traffic.json"), "w") as f:
        json.dump(rec, f, indent=1)
    print(json.dumps(rec))


if __name__ == "__main__":
    main()
