"""End-to-end timing of the Python drop-in with ordinary (pageable) numpy arrays."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from nbodyhpc.kdtree import KDTree

side = int(sys.argv[1]) if len(sys.argv) > 1 else 256
m = int(float(sys.argv[2])) if len(sys.argv) > 2 else 20_000_000
rng = np.random.Generator(np.random.Philox(1))
pts = rng.random((side ** 3, 3), dtype=np.float32)
q = rng.random((m, 3), dtype=np.float32)
for rep in range(3):
    t0 = time.perf_counter(); tree = KDTree(pts, leafsize=64, boxsize=1.0); t1 = time.perf_counter()
    print(f"build from numpy: {t1 - t0:.3f} s ({side**3 / (t1 - t0) / 1e6:.0f} Mpts/s)", flush=True)
for rep in range(6):
    t0 = time.perf_counter(); d, i = tree.query(q, k=8); t1 = time.perf_counter()
    print(f"query {m:.1e} numpy queries k=8: {t1 - t0:.3f} s = {m / (t1 - t0) / 1e6:.0f} Mq/s", flush=True)
    del d, i  # released results are recycled (and page-locked in the background) by the next calls
    time.sleep(1.5 if rep < 2 else 0.0)
q1 = q[:1000].copy()
t0 = time.perf_counter()
for _ in range(100): tree.query(q1, k=8)
print(f"1000-query batches: {(time.perf_counter() - t0) * 10:.3f} ms per call")
q1 = q[:1].copy()
t0 = time.perf_counter()
for _ in range(100): tree.query(q1, k=8)
print(f"single queries: {(time.perf_counter() - t0) * 10:.3f} ms per call")
