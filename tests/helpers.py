"""Shared helpers of the parity tests (test infrastructure; may import oracle/)."""
import numpy as np

from oracle import Oracle, Reference, compare_knn  # noqa: F401


def philox(n, seed, box=1.0):
    """The reference's test fixture generator (kdtree_utils.hpp:16-46), via the C restatement
    (pinned bit-equal to the reference's in tests/test_oracle.py)."""
    return Oracle.philox_points(n, seed, box)


def checker_tree(points, leafsize, boxsize):
    """The strongest checker available: the compiled reference if present, else the restatement."""
    if Reference.available():
        return Reference.Tree(points, leafsize, boxsize)
    return Oracle.Tree(points, leafsize, boxsize)


def leaf_sets(nodes, idx):
    """{(left, right): sorted original indices} for every leaf."""
    out = {}
    for nd in nodes[nodes["dim"] == -1]:
        out[(int(nd["left"]), int(nd["right"]))] = np.sort(idx[nd["left"]:nd["right"]])
    return out


def assert_parity(d, i, d_ref, i_ref, points, queries, boxsize, allow_ties=True):
    rep = compare_knn(d, i, d_ref, i_ref, points, queries, boxsize)
    assert rep.ok, rep
    if not allow_ties:
        assert rep.rows_equal == rep.rows, rep
    return rep
