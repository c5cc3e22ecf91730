"""Host-side checks of bench.py (no GPU): defaults follow the north star (one 10^8-query batch, strong scaling,
configuration 3), the byte model is SURVEY.md 8(d)'s, the traffic record is tied to the kernel source."""
import hashlib
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture()
def bench(monkeypatch):
    monkeypatch.setattr(sys, "argv", ["bench.py"])
    sys.path.insert(0, ROOT)
    import importlib

    import bench as module

    return importlib.reload(module)


def test_defaults_are_the_headline_configuration(bench, monkeypatch):
    args = bench.parse_args()
    assert (args.gpus, args.config, args.scaling, args.side, args.queries, args.k, args.leaf) == \
        (1, 3, "strong", 512, 100_000_000, 8, 64)
    assert bench.metric_name(args) == "kNN queries/sec (k=8, 512^3 periodic tree)"
    assert "of the whole job" in bench.workload_name(args)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--config", "4"])
    args4 = bench.parse_args()
    assert args4.k == 32 and "clustered" in bench.metric_name(args4) and "Zel'dovich" in bench.workload_name(args4)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--config", "4", "-k", "8"])
    with pytest.raises(SystemExit):
        bench.parse_args()


def test_byte_model_is_the_surveys(bench):
    # B_q = 12 + 8k + 16 V_n + 16 V_p; headline calibration 4 715 B
    assert bench.algorithmic_bytes_per_query(8, 34.1, 255.8) == pytest.approx(12 + 64 + 545.6 + 4092.8)
    assert bench.algorithmic_bytes_per_query(32, 45.4, 470.4, rows=False) == pytest.approx(12 + 16 * 45.4 + 16 * 470.4)


def test_traffic_record_matches_the_kernel_source(bench):
    """profiles/knn_traffic.json names the SHA-1 of knn_query.cuh it was captured from; bench.py must say STALE
    exactly when the kernels have changed since (re-capture: scripts/refresh_traffic.py)."""
    with open(os.path.join(ROOT, "profiles", "knn_traffic.json")) as f:
        rec = json.load(f)
    with open(os.path.join(ROOT, "nbodyhpc_b200", "csrc", "knn_query.cuh"), "rb") as f:
        current = hashlib.sha1(f.read()).hexdigest()
    args = bench.parse_args()
    traffic, note = bench.ncu_traffic(args)
    assert traffic == rec["dram_bytes_per_launch"]
    # a capture older than the kernels is reported as such, never passed off as current
    assert note.startswith("STALE") == (rec["kernel_source_sha1"] != current)
    args.k = 4
    assert bench.ncu_traffic(args)[0] is None  # another workload: no figure rather than a wrong one


def test_strong_split_covers_the_batch():
    sys.path.insert(0, ROOT)
    from nbodyhpc_b200.dist import shard_range

    for world in (1, 2, 4, 8):
        pieces = [shard_range(100_000_000, r, world) for r in range(world)]
        assert pieces[0][0] == 0 and pieces[-1][1] == 100_000_000
        assert all(pieces[i][1] == pieces[i + 1][0] for i in range(world - 1))
        assert max(e - b for b, e in pieces) - min(e - b for b, e in pieces) <= world
