"""Pins the oracle (oracle/knn_oracle.c, the plain-C restatement) against
  (1) the committed golden vectors generated from the reference itself (tests/golden/), and
  (2) the compiled reference (oracle/_ref) where it is available (this container),
and replays the reference tests that need no kd-tree product at all (metrics, fixtures)."""
import json
import os

import numpy as np
import pytest

from helpers import Oracle, Reference, compare_knn, philox

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "reference_vectors.npz"))
with open(os.path.join(HERE, "golden", "reference_vectors.json")) as f:
    MANIFEST = json.load(f)

needs_ref = pytest.mark.skipif(not Reference.available(), reason="compiled reference (oracle/_ref) not available")


def test_philox_fixture_generator_matches_golden():
    got = philox(16, 42, 2.0)
    assert np.array_equal(got.view(np.uint32), GOLD["philox/seed42_box2"].view(np.uint32))
    assert (got > 0).all() and (got <= 2.0).all()


@pytest.mark.parametrize("case", MANIFEST, ids=[c["name"] for c in MANIFEST])
def test_oracle_matches_golden_vectors(case):
    name, box = case["name"], case["box"]
    pts = philox(case["n"], case["point_seed"], box or 1.0)
    q = philox(case["m"], case["query_seed"], box or 1.0)
    tree = Oracle.Tree(pts, case["leaf"], box)
    assert tree.n == case["n_padded"] and tree.size == case["n_nodes"]
    assert Oracle.expected_num_nodes(case["n"], case["leaf"]) == case["n_nodes"]
    d, i, stats = tree.query(q, case["k"], return_stats=True)
    rep = compare_knn(d, i, GOLD[f"{name}/d"], GOLD[f"{name}/i"], pts, q, box)
    assert rep.ok and rep.rows_equal + rep.rows_equal_after_tie_canonicalisation == rep.rows, rep
    # exhaustive scan (find_nearest_naive, tests/test.cpp:14-37) agrees as well
    d_bf, i_bf = tree.query(q, case["k"], brute=True)
    rep = compare_knn(d_bf, i_bf, GOLD[f"{name}/d"], GOLD[f"{name}/i"], pts, q, box)
    assert rep.ok, rep
    # traversal counters equal the reference's KDTreeQueryStatistics
    assert np.array_equal(stats, GOLD[f"{name}/stats"])
    # node array: topology always identical; splits identical when no coordinate repeats
    nodes = tree.nodes()
    gold_nodes = GOLD[f"{name}/nodes"].view(nodes.dtype)
    for f_ in ("dim", "left", "right"):
        assert np.array_equal(nodes[f_], gold_nodes[f_])
    if all(len(np.unique(pts[:, a])) == len(pts) for a in range(3)):
        assert np.array_equal(nodes.view(np.uint8), GOLD[f"{name}/nodes"])


def test_metrics_match_golden_known_answers():
    pts = philox(100, 42, 1.0)
    box6 = np.array([0.2, 0.5, 0.4, 0.6, 0.0, 0.1], np.float32)
    q0 = np.array([0.9, 0.05, 0.5], np.float32)
    L = Oracle.lib()
    for key, fn, arg in (("metric/box_periodic", L.orc_box_distance, 1.0), ("metric/box_open", L.orc_box_distance, -1.0)):
        got = np.array([fn(p, box6, arg) for p in pts], np.float32)
        assert np.array_equal(got.view(np.uint32), GOLD[key].view(np.uint32)), key
    for key, arg in (("metric/point_periodic", 1.0), ("metric/point_open", -1.0)):
        got = np.array([L.orc_point_distance(p, q0, arg) for p in pts], np.float32)
        assert np.array_equal(got.view(np.uint32), GOLD[key].view(np.uint32)), key


def test_periodic_box_distance_is_min_over_27_images():
    """tests/test.cpp:116-145 (KDTreeMetric.TestL2PeriodicBox3D)."""
    pts = philox(100, 42, 1.0)
    box6 = np.array([0.2, 0.5, 0.4, 0.6, 0.0, 0.1], np.float32)
    L = Oracle.lib()
    for p in pts:
        d = L.orc_box_distance(p, box6, 1.0)
        naive = min(L.orc_box_distance((p + np.array(s, np.float32)).astype(np.float32), box6, -1.0)
                    for s in np.ndindex(3, 3, 3) for s in [np.array(s) - 1])
        assert abs(d - naive) < 1e-6


def test_padding_and_unfilled_slots():
    """Quirks 1 and 4 of SURVEY.md section 5: n is the padded count; k > N leaves
    (sqrt(FLT_MAX), 0xFFFFFFFF) slots."""
    pts = philox(5, 1)
    tree = Oracle.Tree(pts, 64, None)
    assert tree.n == 8 and tree.size == 1
    d, i = tree.query(philox(4, 2), 8)
    assert (i[:, 5:] == 0xFFFFFFFF).all()
    assert (d[:, 5:] == np.sqrt(np.float32(np.finfo(np.float32).max))).all()
    assert (np.sort(i[:, :5], axis=1) == np.arange(5)).all()
    with pytest.raises(RuntimeError, match="all points must be within the box"):
        Oracle.Tree(pts + 2, 64, 1.0)


@needs_ref
@pytest.mark.parametrize("n,box,leaf", [(10, None, 32), (1000, 2.0, 64), (50_000, None, 64), (50_000, 1.0, 128)])
def test_oracle_matches_compiled_reference(n, box, leaf):
    pts = Reference.philox_points(n, 42, box or 1.0)
    assert np.array_equal(pts, philox(n, 42, box or 1.0))
    q = philox(400, 43, box or 1.0)
    to, tr = Oracle.Tree(pts, leaf, box), Reference.Tree(pts, leaf, box)
    assert (to.n, to.size) == (tr.n, tr.size)
    for f_ in ("dim", "left", "right"):
        assert np.array_equal(to.nodes()[f_], tr.nodes()[f_])
    for k in (1, 4, 8, 20):
        d, i, s = to.query(q, k, return_stats=True)
        d_ref, i_ref, s_ref = tr.query(q, k, workers=2, return_stats=True)
        assert compare_knn(d, i, d_ref, i_ref, pts, q, box).ok
        assert np.array_equal(s, s_ref)
        d_bf, i_bf = to.query(q, k, brute=True)
        assert compare_knn(d_bf, i_bf, d_ref, i_ref, pts, q, box).ok


def test_parity_checker_flags_real_errors_and_accepts_ties():
    pts = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [5, 5, 5]], np.float32)
    q = np.array([[0.5, 0.5, 0]], np.float32)  # points 0,1,2 tie... (0 and 1 and 2 at the same distance)
    d = np.sqrt(np.array([[0.5, 0.5]], np.float32))
    assert compare_knn(d, np.array([[0, 1]], np.uint32), d, np.array([[1, 0]], np.uint32), pts, q).rows_equal_after_tie_canonicalisation == 1
    assert compare_knn(d, np.array([[0, 1]], np.uint32), d, np.array([[2, 1]], np.uint32), pts, q).rows_boundary_tie_verified == 1
    assert compare_knn(d, np.array([[0, 1]], np.uint32), d, np.array([[3, 1]], np.uint32), pts, q).rows_wrong == 1
    assert compare_knn(d, np.array([[0, 1]], np.uint32), d * 2, np.array([[0, 1]], np.uint32), pts, q).rows_wrong == 1
