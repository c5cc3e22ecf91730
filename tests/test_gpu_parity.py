"""GPU parity tests: the CUDA path, called through the C ABI (ctypes) and through the pybind
module, against the oracle on the same seeded inputs.  Bit-exact bar: distances bit-equal,
indices equal (exact ties canonicalised by index as SURVEY.md 8(c) prescribes).

The cases replay the reference's own tests:
  tests/test.cpp:43-111            N in {10,100,1000}, k=4, leaf 32/64, open + periodic box 2.0
  tests/test_inserters.cpp:120-121 {seed,k,n} = {42,1,53} {43,4,79} {44,7,123} {45,13,156} {46,17,179}
  tests/test_builders.cpp:65-127   tuples preserved, every index in exactly one leaf, leaves % 8
  kdtree/tests/test_kdtree.py      scipy comparison, PCG64(42), 10 000 points, k=4
"""
import numpy as np
import pytest

from helpers import (Oracle, Reference, assert_parity, check_tree_invariants, checker_tree, compare_knn,
                     leaf_sets, philox)

pytestmark = pytest.mark.gpu

SQRT_FLT_MAX = np.sqrt(np.float32(np.finfo(np.float32).max))


# ---- the reference's C++ tests ------------------------------------------------------------------
@pytest.mark.parametrize("n", [10, 100, 1000])
@pytest.mark.parametrize("leaf,query", [(32, (0.4, 0.5, 0.6)), (64, (0.5, 0.5, 0.5))])
def test_reference_fixture_open(gpu, n, leaf, query):
    pts = philox(n, 42)
    q = np.array([query], np.float32)
    tree = gpu.Tree.build(pts, leaf)
    d, i = tree.query(q, 4)
    ref = checker_tree(pts, leaf, None)
    d_ref, i_ref = ref.query(q, 4)
    assert_parity(d, i, d_ref, i_ref, pts, q, None, allow_ties=False)
    nodes = tree.nodes()
    leaves = nodes[nodes["dim"] == -1]
    assert ((leaves["right"] - leaves["left"]) % 8 == 0).all()
    assert (np.diff(d[0]) >= 0).all()


@pytest.mark.parametrize("n", [10, 100, 1000])
def test_reference_fixture_periodic(gpu, n):
    box = 2.0
    pts = philox(n, 42, box)
    q = philox(100, 43, box)
    tree = gpu.Tree.build(pts, 64, box)
    d, i = tree.query(q, 4)
    ref = checker_tree(pts, 64, box)
    d_ref, i_ref = ref.query(q, 4)
    assert_parity(d, i, d_ref, i_ref, pts, q, box, allow_ties=False)
    # and against the exhaustive scan (find_nearest_naive)
    d_bf, i_bf = Oracle.Tree(pts, 64, box).query(q, 4, brute=True)
    assert_parity(d, i, d_bf, i_bf, pts, q, box, allow_ties=False)


@pytest.mark.parametrize("seed,k,n", [(42, 1, 53), (43, 4, 79), (44, 7, 123), (45, 13, 156), (46, 17, 179)])
@pytest.mark.parametrize("box", [None, 2.0])
def test_reference_inserter_configs(gpu, seed, k, n, box):
    pts = philox(n, seed, box or 1.0)
    q = philox(64, seed + 100, box or 1.0)
    tree = gpu.Tree.build(pts, 32, box)
    d, i = tree.query(q, k)
    d_ref, i_ref = checker_tree(pts, 32, box).query(q, k)
    assert_parity(d, i, d_ref, i_ref, pts, q, box, allow_ties=False)


# ---- tree structure -----------------------------------------------------------------------------
@pytest.mark.parametrize("n,leaf", [(80, 32), (1000, 16), (4096, 64), (100003, 64), (100003, 128)])
def test_tree_structure_matches_reference(gpu, n, leaf):
    pts = philox(n, 7)
    tree = gpu.Tree.build(pts, leaf)
    ref = checker_tree(pts, leaf, None)
    assert tree.n == ref.n and tree.size == ref.size
    nodes, rnodes = tree.nodes(), ref.nodes()
    # topology is a function of counts only: dim/left/right identical everywhere
    for f in ("dim", "left", "right"):
        assert np.array_equal(nodes[f], rnodes[f]), f
    x, y, z, idx = tree.points()
    # tuples preserved + every index in exactly one leaf (tests/test_builders.cpp:65-80,106-127)
    assert np.array_equal(np.sort(idx), np.arange(tree.n, dtype=np.uint32))
    real = idx < n
    assert np.array_equal(np.stack([x, y, z], 1)[real], pts[idx[real]])
    assert (x[~real] == np.finfo(np.float32).max).all()
    # split invariant: left subtree <= split <= right subtree along dim
    coords = (x, y, z)
    def span(node):
        nd = nodes[node]
        if nd["dim"] == -1:
            return int(nd["left"]), int(nd["right"])
        l0, _ = span(nd["left"]); _, r1 = span(nd["right"])
        return l0, r1
    for k_node in np.nonzero(nodes["dim"] >= 0)[0][:200]:
        nd = nodes[k_node]
        l0, l1 = span(nd["left"]); r0, r1 = span(nd["right"])
        c = coords[nd["dim"]]
        assert c[l0:l1].max() <= nd["split"] <= c[r0:r1].min()
        assert nd["split"] == c[r0:r1].min()  # the median element is the smallest of the right child
    # where no coordinate value repeats, the split values (and leaf memberships) are unique too
    if all(len(np.unique(pts[:, a])) == n for a in range(3)):
        assert np.array_equal(nodes["split"], rnodes["split"])
        rx, ry, rz, ridx = ref.points()
        assert leaf_sets(nodes, idx).keys() == leaf_sets(rnodes, ridx).keys()
        for key, val in leaf_sets(nodes, idx).items():
            assert np.array_equal(val, leaf_sets(rnodes, ridx)[key])


def _tie_heavy_sets():
    rng = np.random.Generator(np.random.Philox(11))
    g = (np.arange(32, dtype=np.float32) + 0.5) / 32
    lattice = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    blobs = np.concatenate([rng.normal(c, s, (40000, 3)) for c, s in [(0.2, 1e-3), (0.7, 1e-2), (0.5, 1e-5)]]
                           + [rng.random((20000, 3))]).astype(np.float32)
    return {
        "lattice32": lattice,                                          # every coordinate takes 32 values
        "lattice32-shuffled": lattice[rng.permutation(len(lattice))],
        "identical": np.full((30000, 3), 0.25, np.float32),             # one value: selection by id alone
        "clustered": np.clip(blobs, 0, 1),                              # 40000 points within 1e-4 of a centre
        "wide-negative": ((rng.random((150000, 3)) - 0.5) * 2e6).astype(np.float32),
        "one-million": philox(1 << 20, 9),                              # 2^20: repeated coordinates straddle medians
    }


@pytest.mark.parametrize("name", ["lattice32", "lattice32-shuffled", "identical", "clustered", "wide-negative",
                                  "one-million"])
@pytest.mark.parametrize("leaf", [16, 64])
def test_build_invariants_under_ties_and_clustering(gpu, name, leaf):
    """Select-and-partition build (tree_topdown.cuh / tree_bottom.cuh) on inputs where coordinates
    repeat massively: topology from counts alone, exact rank-median splits, tuples preserved, and the
    kNN answer equal to the reference's (tests/test_builders.cpp:65-127 strengthened)."""
    pts = _tie_heavy_sets()[name]
    n = len(pts)
    tree = gpu.Tree.build(pts, leaf)
    nodes = tree.nodes()
    plan = gpu.plan_topology(n, leaf, 8)[0]
    for f in ("dim", "left", "right"):
        assert np.array_equal(nodes[f], plan[f]), f
    x, y, z, idx = tree.points()
    assert np.array_equal(np.sort(idx), np.arange(tree.n, dtype=np.uint32))
    real = idx < n
    assert np.array_equal(np.stack([x, y, z], 1)[real], pts[idx[real]])
    check_tree_invariants(nodes, x, y, z, sample_every=1 if len(nodes) < 20000 else 5)
    q = philox(500, 77)
    d, i = tree.query(q, 8)
    d_ref, i_ref = checker_tree(pts, leaf, None).query(q, 8, workers=0)
    assert_parity(d, i, d_ref, i_ref, pts, q, None)


@pytest.mark.parametrize("n,leaf,block", [(70000, 9000, 8), (200000, 20000, 8), (100000, 64, 16), (100000, 100, 64),
                                          (8192, 16, 8), (8200, 64, 8), (3_000_001, 64, 8)])
def test_build_shapes(gpu, n, leaf, block):
    """Leaves larger than the shared-memory sub-tree limit, other block sizes, segment sizes at and
    around the bottom kernel's capacity, and a tree whose top phase runs nine levels."""
    pts = philox(n, 5)
    tree = gpu.Tree.build(pts, leaf, block_size=block)
    nodes = tree.nodes()
    plan = gpu.plan_topology(n, leaf, block)[0]
    for f in ("dim", "left", "right"):
        assert np.array_equal(nodes[f], plan[f]), f
    x, y, z, idx = tree.points()
    assert np.array_equal(np.sort(idx), np.arange(tree.n, dtype=np.uint32))
    check_tree_invariants(nodes, x, y, z, block=block, sample_every=1 if len(nodes) < 20000 else 7)
    if block == 8:
        ref = checker_tree(pts, leaf, None)
        assert tree.size == ref.size and tree.n == ref.n
        q = philox(300, 78)
        d, i = tree.query(q, 4)
        d_ref, i_ref = ref.query(q, 4, workers=0)
        assert_parity(d, i, d_ref, i_ref, pts, q, None)


def test_build_is_deterministic(gpu):
    """The order in which the partition passes write a segment varies from run to run; the tree may
    not: same node array and same point order, also under heavy ties."""
    for pts in (philox(300000, 21), _tie_heavy_sets()["lattice32-shuffled"]):
        a, b = gpu.Tree.build(pts, 32), gpu.Tree.build(pts, 32)
        assert np.array_equal(a.nodes(), b.nodes())
        for u, v in zip(a.points(), b.points()):
            assert np.array_equal(u, v)


def test_build_soa_with_custom_indices(gpu):
    n = 2048
    pts = philox(n, 3)
    custom = (np.arange(n, dtype=np.uint32) * 7 + 11)
    tree = gpu.Tree.build_soa(pts[:, 0], pts[:, 1], pts[:, 2], custom, 32)
    q = philox(200, 4)
    d, i = tree.query(q, 5)
    d_ref, i_ref = checker_tree(pts, 32, None).query(q, 5)
    assert np.array_equal(d.view(np.uint32), d_ref.view(np.uint32))
    assert np.array_equal(i, custom[i_ref])
    with pytest.raises(gpu.NbkError, match="block_size must divide the number of points."):
        gpu.Tree.build_soa(pts[:100, 0], pts[:100, 1], pts[:100, 2], custom[:100], 32)


# ---- batched queries at CPU-test scale ------------------------------------------------------------
@pytest.mark.parametrize("box", [None, 1.0])
@pytest.mark.parametrize("k", [1, 2, 3, 8, 16, 17, 32, 64])
def test_batched_query_vs_checker(gpu, box, k):
    n, m = 200_000, 20_000
    pts, q = philox(n, 42, box or 1.0), philox(m, 43, box or 1.0)
    tree = gpu.Tree.build(pts, 64, box)
    d, i = tree.query(q, k)
    d_ref, i_ref = checker_tree(pts, 64, box).query(q, k, workers=0)
    rep = assert_parity(d, i, d_ref, i_ref, pts, q, box)
    assert rep.rows_equal >= rep.rows - 5


def test_config1_million_periodic_k8(gpu):
    """BASELINE.json configs[0]: 1M uniform points in the periodic unit box, k=8, 1M queries
    (checked against the reference on a 100k-query sample, and fully against invariants)."""
    n = m = 1_000_000
    pts, q = philox(n, 42), philox(m, 43)
    tree = gpu.Tree.build(pts, 64, 1.0)
    d, i = tree.query(q, 8)
    sample = slice(0, 100_000)
    d_ref, i_ref = checker_tree(pts, 64, 1.0).query(q[sample], 8, workers=0)
    assert_parity(d[sample], i[sample], d_ref, i_ref, pts, q[sample], 1.0)
    assert (np.diff(d, axis=1) >= 0).all()
    assert (i < n).all()


def test_self_query_open_config2_shape(gpu):
    """BASELINE.json configs[1] at reduced size: non-periodic self-query, k=1..16."""
    n = 64 ** 3
    pts = philox(n, 5)
    tree = gpu.Tree.build(pts, 64)
    ref = checker_tree(pts, 64, None)
    for k in (1, 2, 4, 8, 16):
        d, i = tree.query(pts, k)
        assert (d[:, 0] == 0).all()
        # a point's nearest neighbour at distance 0 is itself unless an exact duplicate precedes it
        assert (pts[i[:, 0]] == pts).all()
        d_ref, i_ref = ref.query(pts[:20000], k, workers=0)
        assert_parity(d[:20000], i[:20000], d_ref, i_ref, pts, pts[:20000], None)


# ---- edge cases -----------------------------------------------------------------------------------
def test_fewer_points_than_k(gpu):
    pts = philox(5, 1)
    tree = gpu.Tree.build(pts, 64)
    q = philox(40, 2)
    d, i = tree.query(q, 8)
    d_ref, i_ref = checker_tree(pts, 64, None).query(q, 8)
    assert np.array_equal(d.view(np.uint32), d_ref.view(np.uint32))
    assert np.array_equal(i, i_ref)
    assert (i[:, 5:] == 0xFFFFFFFF).all() and (d[:, 5:] == SQRT_FLT_MAX).all()


def test_empty_inputs(gpu):
    tree = gpu.Tree.build(np.zeros((0, 3), np.float32), 64)
    assert tree.n == 0 and tree.size == 1
    d, i = tree.query(philox(3, 1), 2)
    assert (i == 0xFFFFFFFF).all() and (d == SQRT_FLT_MAX).all()
    tree = gpu.Tree.build(philox(100, 1), 64)
    d, i = tree.query(np.zeros((0, 3), np.float32), 3)
    assert d.shape == (0, 3) and i.shape == (0, 3)


def test_exact_ties_ordered_by_index(gpu):
    """All-identical points and lattice points: every distance ties; the product orders ties by
    index (north star), the multiset of distances equals the reference's."""
    pts = np.tile(np.array([[0.25, 0.5, 0.75]], np.float32), (300, 1))
    tree = gpu.Tree.build(pts, 16)
    q = np.array([[0.1, 0.2, 0.3], [0.25, 0.5, 0.75]], np.float32)
    d, i = tree.query(q, 8)
    assert np.array_equal(i, np.tile(np.arange(8, dtype=np.uint32), (2, 1)))
    g = np.arange(8, dtype=np.float32) / 8
    lattice = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    for box in (None, 1.0):
        tree = gpu.Tree.build(lattice, 16, box)
        ql = lattice + np.float32(1 / 16)
        d, i = tree.query(ql, 8)
        d_bf, i_bf = Oracle.Tree(lattice, 16, box).query(ql, 8, brute=True)
        assert np.array_equal(d.view(np.uint32), d_bf.view(np.uint32))
        assert np.array_equal(i, i_bf)
        d_ref, i_ref = checker_tree(lattice, 16, box).query(ql, 8)
        rep = compare_knn(d, i, d_ref, i_ref, lattice, ql, box)
        assert rep.ok, rep


def test_periodic_wrap_and_out_of_box_queries(gpu):
    box = 1.0
    pts = philox(50_000, 9)
    tree = gpu.Tree.build(pts, 64, box)
    rng = np.random.default_rng(0)
    # queries hugging the faces/corners, where wrapped images matter
    q = rng.uniform(0, 1, (4000, 3)).astype(np.float32)
    q[:2000] = np.where(rng.random((2000, 3)) < 0.5, q[:2000] * 1e-3, 1 - q[:2000] * 1e-3).astype(np.float32)
    d, i = tree.query(q, 8)
    d_ref, i_ref = checker_tree(pts, 64, box).query(q, 8)
    assert_parity(d, i, d_ref, i_ref, pts, q, box)
    # queries outside the box are not validated by the reference (pybind.cpp:35-47 checks points
    # only); the product returns the exact 3-image answer = the exhaustive scan
    q_out = (rng.uniform(-0.4, 1.4, (2000, 3))).astype(np.float32)
    d, i = tree.query(q_out, 4)
    d_bf, i_bf = Oracle.Tree(pts, 64, box).query(q_out, 4, brute=True)
    assert np.array_equal(d.view(np.uint32), d_bf.view(np.uint32))
    assert np.array_equal(i, i_bf)


def test_points_on_box_faces_and_validation(gpu):
    box = 2.0
    pts = philox(1000, 11, box)
    pts[0] = (0.0, 0.0, 0.0)
    pts[1] = (box, box, box)  # inclusive upper bound is legal (pybind.cpp:39)
    tree = gpu.Tree.build(pts, 32, box)
    q = philox(300, 12, box)
    d, i = tree.query(q, 3)
    d_ref, i_ref = checker_tree(pts, 32, box).query(q, 3)
    assert_parity(d, i, d_ref, i_ref, pts, q, box)
    bad = pts.copy()
    bad[17, 1] = np.nextafter(np.float32(box), np.float32(10))
    with pytest.raises(gpu.NbkError, match="all points must be within the box"):
        gpu.Tree.build(bad, 32, box)
    bad[17, 1] = -1e-6
    with pytest.raises(gpu.NbkError, match="all points must be within the box"):
        gpu.Tree.build(bad, 32, box)
    gpu.Tree.build(bad, 32, None)  # open boundaries accept anything


def test_clustered_points(gpu):
    """Strongly non-uniform input (config 4's regime): pruning bounds must stay exact."""
    rng = np.random.default_rng(3)
    centres = rng.uniform(0, 1, (20, 3))
    pts = (centres[rng.integers(0, 20, 150_000)] + rng.normal(0, 0.004, (150_000, 3))) % 1.0
    pts = pts.astype(np.float32)
    pts = np.clip(pts, 0, 1)
    q = rng.uniform(0, 1, (20_000, 3)).astype(np.float32)
    for box in (None, 1.0):
        tree = gpu.Tree.build(pts, 64, box)
        d, i = tree.query(q, 8)
        d_ref, i_ref = checker_tree(pts, 64, box).query(q, 8, workers=0)
        assert_parity(d, i, d_ref, i_ref, pts, q, box)


def test_metric_override_per_call(gpu):
    """KDTree::find_closest<Distance> picks the metric per call (kdtree.hpp:207-210)."""
    pts, q = philox(30_000, 21), philox(1000, 22)
    tree = gpu.Tree.build(pts, 64)  # built open
    d, i = tree.query(q, 4, periodic=1, boxsize=1.0)
    d_ref, i_ref = checker_tree(pts, 64, 1.0).query(q, 4)
    assert_parity(d, i, d_ref, i_ref, pts, q, 1.0)


def test_statistics_match_reference(gpu):
    """KDTreeQueryStatistics on the device == the reference's counters (exactly, when no coordinate
    repeats so that both trees hold the same leaves)."""
    n = 4000
    for seed in range(31, 200):  # first fixture without a repeated coordinate value
        pts = philox(n, seed)
        if all(len(np.unique(pts[:, a])) == n for a in range(3)):
            break
    else:
        pytest.fail("no duplicate-free fixture found")
    q = philox(500, 32)
    for box in (None, 1.0):
        tree = gpu.Tree.build(pts, 32, box)
        ref = checker_tree(pts, 32, box)
        for k in (1, 4, 8):
            _, _, s_ref = ref.query(q, k, return_stats=True)
            assert np.array_equal(tree.stats(q, k), s_ref)


# ---- the Python drop-in (pybind module) -----------------------------------------------------------
def test_python_api_matches_scipy_like_the_reference_tests(gpu):
    import scipy.spatial
    from nbodyhpc.kdtree import KDTree  # the reference's import path

    rng = np.random.Generator(np.random.PCG64(42))
    points = rng.uniform(0, 1, size=(10000, 3))
    query_points = rng.uniform(0, 1, size=(200, 3))
    tree = KDTree(points)
    d_ref, i_ref = scipy.spatial.KDTree(points).query(query_points, k=4)
    d, i = tree.query(query_points, k=4)
    assert d.dtype == np.float32 and i.dtype == np.uint32 and d.shape == (200, 4)
    assert np.allclose(d_ref, d) and np.all(i_ref == i)
    assert tree.n == 10000 and tree.size == Oracle.expected_num_nodes(10000, 128) and not tree.periodic
    assert tree.boxsize == 0.0

    boxsize = 2.0
    points = rng.uniform(0, boxsize, size=(10000, 3)).astype(np.float32)
    query_points = rng.uniform(0, boxsize, size=(200, 3)).astype(np.float32)
    tree = KDTree(points, boxsize=boxsize)
    d_ref, i_ref = scipy.spatial.KDTree(points, boxsize=boxsize).query(query_points, k=4)
    d, i = tree.query(query_points, k=4)
    assert np.allclose(d_ref, d) and np.all(i_ref == i)
    assert tree.periodic and tree.boxsize == 2.0


def test_python_api_behaviour(gpu):
    from nbodyhpc_b200.kdtree import KDTree

    pts = philox(1001, 1)
    with pytest.warns(UserWarning, match="Unrecognized keyword arguments"):
        tree = KDTree(pts, leafsize=64, balanced_tree=False)
    assert tree.n == 1008  # padded count (pybind.cpp:71)
    with pytest.raises(RuntimeError, match="k must be positive integer"):
        tree.query(pts[:3], k=0)
    with pytest.raises(RuntimeError, match=r"positions must be a 2D array of shape \(N, 3\)"):
        KDTree(np.zeros((5, 2)))
    with pytest.raises(RuntimeError, match="all points must be within the box"):
        KDTree(pts + 5, boxsize=1.0)
    # N-d query shapes (broken in the reference, __init__.py:52-54; fixed here)
    d, i = tree.query(pts[:24].reshape(2, 3, 4, 3), k=2)
    assert d.shape == (2, 3, 4, 2) and i.shape == (2, 3, 4, 2)
    d2, i2 = tree.query(pts[:24], k=2, workers=-1)
    assert np.array_equal(d.reshape(-1, 2), d2) and np.array_equal(i.reshape(-1, 2), i2)
    # float64 input is cast to float32 like py::array_t<float> does (pybind.cpp:77,91)
    d3, _ = tree.query(pts[:24].astype(np.float64), k=2)
    assert np.array_equal(d3, d2)
    nodes = tree.nodes()
    assert len(nodes) == tree.size and nodes[0]["dim"] == 0


@pytest.mark.parametrize("n", [0, 1, 7, 8, 9, 15, 16, 17, 24, 100])
def test_tiny_trees(gpu, n):
    """Point counts around the padding unit (8) and the minimum leaf (16): structure and answers."""
    pts = philox(n, 31) if n else np.zeros((0, 3), np.float32)
    tree = gpu.Tree.build(pts, 16)
    nodes = tree.nodes()
    x, y, z, idx = tree.points()
    check_tree_invariants(nodes, x, y, z)
    q = philox(20, 32)
    k = 3
    d, i = tree.query(q, k)
    if n == 0:
        # (the reference itself crashes when an empty tree is queried; the drop-in reports no neighbours)
        assert tree.n == 0 and tree.size == 1
        assert (d == SQRT_FLT_MAX).all() and (i == 0xFFFFFFFF).all()
        return
    ref = checker_tree(pts, 16, None)
    assert tree.n == ref.n and tree.size == ref.size
    d_ref, i_ref = ref.query(q, k)
    assert np.array_equal(d.view(np.uint32), d_ref.view(np.uint32)) and np.array_equal(i, i_ref)


def test_non_finite_coordinates_do_not_break_the_build(gpu):
    """NaN / inf coordinates have no defined behaviour in the reference (comparisons are simply false);
    here they must neither hang nor crash the selection, and the finite points stay searchable."""
    rng = np.random.Generator(np.random.Philox(3))
    pts = rng.random((50_000, 3), dtype=np.float32)
    pts[::97, 0] = np.nan
    pts[5::101, 1] = np.inf
    pts[7::103, 2] = -np.inf
    tree = gpu.Tree.build(pts, 32)
    x, y, z, idx = tree.points()
    assert np.array_equal(np.sort(idx), np.arange(tree.n, dtype=np.uint32))
    q = rng.random((200, 3), dtype=np.float32)
    d, i = tree.query(q, 4)
    finite = np.isfinite(pts).all(1)
    assert finite[i].all()
    # the neighbours returned are real distances to finite points (no pruning guarantees with NaN splits)
    recomputed = np.sqrt(((pts[i] - q[:, None, :]) ** 2).sum(-1, dtype=np.float32))
    assert np.allclose(recomputed, d, rtol=1e-6)


def test_sort_build_cross_check(gpu, tmp_path):
    """NBK_BUILD=sort (one segmented radix sort per level, the first implementation) and the default
    select-and-partition build are independent code paths: where no coordinate repeats they must give
    the same node array and the same point set in every leaf."""
    import os
    import subprocess
    import sys

    n = 200_003
    rng = np.random.Generator(np.random.Philox(17))
    # every coordinate value occurs once per axis
    pts = np.stack([(rng.permutation(n) + 0.5) / n for _ in range(3)], 1).astype(np.float32)
    assert all(len(np.unique(pts[:, a])) == n for a in range(3))
    np.save(tmp_path / "pts.npy", pts)
    script = (
        "import sys, numpy as np; sys.path.insert(0, %r)\n"
        "from nbodyhpc_b200 import capi\n"
        "t = capi.Tree.build(np.load(%r), 32)\n"
        "x, y, z, idx = t.points()\n"
        "np.savez(%r, nodes=t.nodes().view(np.uint8), idx=idx)\n"
    )
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = {}
    for mode in ("select", "sort"):
        dst = str(tmp_path / f"{mode}.npz")
        env = dict(os.environ, NBK_BUILD=mode)
        subprocess.run([sys.executable, "-c", script % (root, str(tmp_path / "pts.npy"), dst)], check=True, env=env,
                       timeout=120)
        out[mode] = np.load(dst)
    assert np.array_equal(out["select"]["nodes"], out["sort"]["nodes"])
    from nbodyhpc_b200.capi import NODE_DTYPE

    nodes = out["select"]["nodes"].view(NODE_DTYPE)
    a, b = leaf_sets(nodes, out["select"]["idx"]), leaf_sets(nodes, out["sort"]["idx"])
    assert a.keys() == b.keys() and all(np.array_equal(a[key], b[key]) for key in a)
