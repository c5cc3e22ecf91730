"""Generates tests/golden/reference_vectors.npz from the REFERENCE ITSELF (oracle/_ref/libnbref.so =
the reference's unmodified sources, see oracle/Makefile).  Run in the build container, where
/root/reference exists:   PYTHONPATH=. python tests/golden/make_golden.py

Inputs are the reference's own Philox fixtures (kdtree_utils.hpp:16-46), so only seeds are stored
for them; outputs (distances, indices, node arrays, counters) are stored verbatim.
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import Reference  # noqa: E402

CASES = [
    # name, n, point seed, m, query seed, box, leaf, k   (tests/test.cpp:43-111 and test_inserters.cpp:120-121)
    ("open_n10", 10, 42, 50, 43, None, 32, 4),
    ("open_n100", 100, 42, 50, 43, None, 32, 4),
    ("open_n1000", 1000, 42, 100, 43, None, 32, 4),
    ("open_n1000_leaf64", 1000, 42, 100, 43, None, 64, 4),
    ("periodic_n10", 10, 42, 100, 43, 2.0, 64, 4),
    ("periodic_n100", 100, 42, 100, 43, 2.0, 64, 4),
    ("periodic_n1000", 1000, 42, 100, 43, 2.0, 64, 4),
    ("ins_42_1_53", 53, 42, 32, 142, None, 32, 1),
    ("ins_43_4_79", 79, 43, 32, 143, 2.0, 32, 4),
    ("ins_44_7_123", 123, 44, 32, 144, None, 32, 7),
    ("ins_45_13_156", 156, 45, 32, 145, 2.0, 32, 13),
    ("ins_46_17_179", 179, 46, 32, 146, None, 32, 17),
    ("open_n20000_k8", 20000, 7, 300, 8, None, 64, 8),
    ("periodic_n20000_k8", 20000, 7, 300, 8, 1.0, 64, 8),
    ("periodic_n20000_k64", 20000, 7, 100, 9, 1.0, 128, 64),
]


def main():
    out = {}
    manifest = []
    for name, n, ps, m, qs, box, leaf, k in CASES:
        pts = Reference.philox_points(n, ps, box or 1.0)
        q = Reference.philox_points(m, qs, box or 1.0)
        tree = Reference.Tree(pts, leaf, box)
        d, i, stats = tree.query(q, k, return_stats=True)
        out[f"{name}/d"] = d
        out[f"{name}/i"] = i
        out[f"{name}/stats"] = stats
        out[f"{name}/nodes"] = tree.nodes().view(np.uint8)
        manifest.append(dict(name=name, n=n, point_seed=ps, m=m, query_seed=qs, box=box, leaf=leaf, k=k,
                             n_padded=tree.n, n_nodes=tree.size))
    # a few raw Philox values pin the fixture generator itself
    out["philox/seed42_box2"] = Reference.philox_points(16, 42, 2.0)
    # metric known answers (kdtree.hpp:20-121) on fixed inputs
    pts = Reference.philox_points(100, 42, 1.0)
    box6 = np.array([0.2, 0.5, 0.4, 0.6, 0.0, 0.1], np.float32)  # tests/test.cpp:118
    L = Reference.lib()
    out["metric/box_periodic"] = np.array([L.ref_box_distance(p, box6, 1.0) for p in pts], np.float32)
    out["metric/box_open"] = np.array([L.ref_box_distance(p, box6, -1.0) for p in pts], np.float32)
    q0 = np.array([0.9, 0.05, 0.5], np.float32)
    out["metric/point_periodic"] = np.array([L.ref_point_distance(p, q0, 1.0) for p in pts], np.float32)
    out["metric/point_open"] = np.array([L.ref_point_distance(p, q0, -1.0) for p in pts], np.float32)
    here = os.path.dirname(os.path.abspath(__file__))
    np.savez_compressed(os.path.join(here, "reference_vectors.npz"), **out)
    with open(os.path.join(here, "reference_vectors.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
