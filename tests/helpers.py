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


def check_tree_invariants(nodes, x, y, z, block=8, sample_every=1):
    """The defining property of the reference's tree (kdtree_impl.hpp:101-125), valid under ties too:
    nodes in pre-order; the left child of an internal node holds ((count/2)/block)*block points, all
    of them <= split <= all points of the right child along dim, and split == min(right child)."""
    coords = (x, y, z)
    beg = np.zeros(len(nodes), np.int64)
    end = np.zeros(len(nodes), np.int64)
    for j in range(len(nodes) - 1, -1, -1):
        nd = nodes[j]
        if nd["dim"] == -1:
            beg[j], end[j] = nd["left"], nd["right"]
            assert (end[j] - beg[j]) % 8 == 0
            continue
        l, r = int(nd["left"]), int(nd["right"])
        assert l == j + 1 and r > l, (j, l, r)
        beg[j], end[j] = beg[l], end[r]
        assert end[l] == beg[r]
        cnt = end[j] - beg[j]
        assert end[l] - beg[l] == (cnt // 2 // block) * block, (j, cnt)
        if j % sample_every:
            continue
        c = coords[nd["dim"]]
        lmax, rmin = c[beg[l]:end[l]].max(), c[beg[r]:end[r]].min()
        assert lmax <= nd["split"] == rmin, (j, int(nd["dim"]), float(lmax), float(nd["split"]), float(rmin))
    assert beg[0] == 0 and end[0] == len(x)
