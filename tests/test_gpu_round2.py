"""GPU tests of the round-2 surface: squared distances, any k (k > 64), the leaf scan + top-k container in
isolation, resumed boundary pass, per-device state, staging of pageable buffers per device."""
import numpy as np
import pytest

from helpers import Oracle, assert_parity, checker_tree, compare_knn, philox

pytestmark = pytest.mark.gpu


# ---- the leaf scan + top-k in isolation (reference: tests/test_asm.cpp:97-199, test_inserters.cpp:159-220) ---
@pytest.mark.parametrize("box", [None, 1.1])
@pytest.mark.parametrize("k", [1, 3, 7, 8, 16, 20, 64, 70])
def test_flat_block_scan_matches_flat_oracle(gpu, k, box):
    """One flat block of n points, no tree: the production scan_leaf + container against the oracle's
    exhaustive scan, n in {8,16,32,256} x 16 seeds, open and periodic (box 1.1 like the reference's asm tests).
    Rows are compared as SQUARED distances, bit for bit, indices exactly."""
    for n in (8, 16, 32, 256):
        for seed in range(16):
            pts = philox(n, 1000 + seed)
            q = philox(24, 2000 + seed)
            if box is not None:
                q = q * np.float32(box)  # queries anywhere in the box; points in [0, 1) subset of it
            ids = (np.arange(n, dtype=np.uint32) * 3 + seed)  # not the identity: the payload is what comes back
            d, i = gpu.scan_block(pts[:, 0], pts[:, 1], pts[:, 2], ids, q, k, boxsize=box, squared=True)
            d_ref, i_ref = Oracle.Tree(pts, 16, box).query(q, k, brute=True, squared=True)
            found = i_ref != 0xFFFFFFFF
            i_ref = np.where(found, ids[np.minimum(i_ref, n - 1)], 0xFFFFFFFF).astype(np.uint32)
            assert np.array_equal(d.view(np.uint32), d_ref.view(np.uint32)), (n, seed)
            assert np.array_equal(i, i_ref), (n, seed)


def test_flat_block_scan_argument_errors(gpu):
    pts, q = philox(12, 1), philox(2, 2)
    ids = np.arange(12, dtype=np.uint32)
    with pytest.raises(gpu.NbkError, match="block_size must be a multiple of 8."):
        gpu.scan_block(pts[:, 0], pts[:, 1], pts[:, 2], ids, q, 2)
    with pytest.raises(gpu.NbkError, match="k must be positive integer"):
        gpu.scan_block(pts[:8, 0], pts[:8, 1], pts[:8, 2], ids[:8], q, 0)


# ---- squared distances ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("box", [None, 1.0])
@pytest.mark.parametrize("k", [1, 5, 8, 24])
def test_squared_distances_bit_exact(gpu, box, k):
    """NBK_QUERY_SQUARED rows == the reference Distance functor's values before postprocess()."""
    n, m = 150_000, 30_000
    pts, q = philox(n, 42), philox(m, 43)
    tree = gpu.Tree.build(pts, 64, box)
    d2, i2 = tree.query(q, k, squared=True)
    d2_ref, i_ref = checker_tree(pts, 64, box).query(q, k, workers=0, squared=True)
    rep = compare_knn(d2, i2, d2_ref, i_ref, pts, q, box, squared=True)
    assert rep.ok, rep
    d, i = tree.query(q, k)
    assert np.array_equal(d, np.sqrt(d2)) and np.array_equal(i, i2)


def test_python_return_squared(gpu):
    from nbodyhpc.kdtree import KDTree

    pts, q = philox(20_000, 3), philox(500, 4)
    tree = KDTree(pts, leafsize=64, boxsize=1.0)
    d, i = tree.query(q, k=6)
    d2, i2 = tree.query(q, k=6, return_squared=True)
    assert d2.dtype == np.float32 and np.array_equal(i, i2) and np.array_equal(np.sqrt(d2), d)
    d2_ref, _ = checker_tree(pts, 64, 1.0).query(q, 6, squared=True)
    assert np.array_equal(d2.view(np.uint32), d2_ref.view(np.uint32))


# ---- any k ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("box", [None, 1.0])
@pytest.mark.parametrize("k", [65, 100, 256])
def test_k_above_64(gpu, box, k):
    """The reference's queue accepts every k (kdtree.cpp:133-141, tournament_tree.hpp:70-77); so does the
    drop-in: above 64 the candidates live in a device-memory heap."""
    n, m = 100_000, 3_000
    pts, q = philox(n, 42), philox(m, 43)
    tree = gpu.Tree.build(pts, 64, box)
    d, i = tree.query(q, k, squared=True)
    d_ref, i_ref = checker_tree(pts, 64, box).query(q, k, workers=0, squared=True)
    rep = compare_knn(d, i, d_ref, i_ref, pts, q, box, squared=True)
    assert rep.ok, rep
    assert rep.rows_equal >= rep.rows - 5


def test_k_above_64_more_than_points_and_batches(gpu, monkeypatch):
    pts, q = philox(90, 5), philox(50, 6)
    tree = gpu.Tree.build(pts, 16, 1.0)
    d, i = tree.query(q, 130)
    d_ref, i_ref = checker_tree(pts, 16, 1.0).query(q, 130)
    assert np.array_equal(d.view(np.uint32), d_ref.view(np.uint32)) and np.array_equal(i, i_ref)
    assert (i[:, 90:] == 0xFFFFFFFF).all()
    # more queries than one global-heap batch holds (k = 600 -> the smallest batch, 75 776 columns)
    pts, q = philox(30_000, 7), philox(80_000, 8)
    tree = gpu.Tree.build(pts, 64, None)
    d, i = tree.query(q, 600)
    sample = np.r_[0:300, 75_700:75_900, 79_800:80_000]
    d_ref, i_ref = checker_tree(pts, 64, None).query(q[sample], 600, workers=0)
    assert_parity(d[sample], i[sample], d_ref, i_ref, pts, q[sample], None)
    assert (np.diff(d, axis=1) >= 0).all()


def test_statistics_and_cdf_above_64(gpu):
    n = 6000
    for seed in range(31, 200):  # first fixture without a repeated coordinate value
        pts = philox(n, seed)
        if all(len(np.unique(pts[:, a])) == n for a in range(3)):
            break
    q = philox(400, 32)
    tree, ref = gpu.Tree.build(pts, 32, 1.0), checker_tree(pts, 32, 1.0)
    for k in (65, 128):
        d_ref, _, s_ref = ref.query(q, k, return_stats=True)
        assert np.array_equal(tree.stats(q, k), s_ref)
    edges = np.linspace(0, 0.5, 33).astype(np.float32)
    counts = tree.knn_cdf(q, [100, 3, 128, 64], edges)
    d_ref, _ = ref.query(q, 128)
    for r, k in enumerate([100, 3, 128, 64]):
        assert np.array_equal(counts[r], np.histogram(d_ref[:, k - 1], edges)[0].astype(np.uint64)), k


# ---- k between the container sizes prunes with the k-th distance ----------------------------------------------------
@pytest.mark.parametrize("k", [3, 5, 6, 7, 9, 12, 33])
def test_odd_k_rows_and_boundary_pass(gpu, k):
    """k that is not a container size (register lists of 1/2/4/8, heaps): exact rows, also for the queries
    whose search continues through the box faces (face-hugging queries resume from the first pass's row)."""
    pts = philox(80_000, 9)
    rng = np.random.default_rng(k)
    q = rng.uniform(0, 1, (6000, 3)).astype(np.float32)
    q[:4000] = np.where(rng.random((4000, 3)) < 0.5, q[:4000] * 2e-3, 1 - q[:4000] * 2e-3).astype(np.float32)
    for box in (None, 1.0):
        tree = gpu.Tree.build(pts, 64, box)
        d, i = tree.query(q, k, squared=True)
        d_ref, i_ref = checker_tree(pts, 64, box).query(q, k, workers=0, squared=True)
        rep = compare_knn(d, i, d_ref, i_ref, pts, q, box, squared=True)
        assert rep.ok, rep
        if box is not None:
            edges = np.linspace(0, 0.2, 21).astype(np.float32)
            counts = tree.knn_cdf(q, [k], edges)  # the from-scratch boundary pass (no rows to resume from)
            assert np.array_equal(counts[0], np.histogram(np.sqrt(d_ref[:, k - 1]), edges)[0].astype(np.uint64))


# ---- per-device state ---------------------------------------------------------------------------------------------------
def test_build_and_query_on_a_second_device(gpu):
    """Function attributes (the bottom kernel's 215 KB of shared memory, the k = 64 heap) are per device."""
    if gpu.lib().nbk_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    pts, q = philox(300_000, 42), philox(5000, 43)
    a = gpu.Tree.build(pts, 64, 1.0, device=0)
    b = gpu.Tree.build(pts, 64, 1.0, device=1)  # built on device 1 after device 0 in one process
    assert a.device == 0 and b.device == 1
    assert np.array_equal(a.nodes(), b.nodes())
    for k in (8, 64, 80):
        da, ia = a.query(q, k)
        db, ib = b.query(q, k)
        assert np.array_equal(da, db) and np.array_equal(ia, ib)


def test_sharded_numpy_queries_are_all_staged(gpu):
    """KDTree(devices=[...]) answers one chunk per replica from host threads; every replica must take the
    pinned staging path for its pageable chunk (one ring per device), none the plain pageable copy."""
    if gpu.lib().nbk_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from nbodyhpc_b200.kdtree import KDTree

    ndev = min(gpu.lib().nbk_device_count(), 4)
    pts = philox(400_000, 42)
    q = np.random.default_rng(1).random((ndev * 3_000_000, 3), dtype=np.float32)
    tree = KDTree(pts, leafsize=64, boxsize=1.0, devices=list(range(ndev)))
    before = gpu.host_path_stats()
    d, i = tree.query(q, k=8)
    after = gpu.host_path_stats()
    assert after["staged_downloads"] - before["staged_downloads"] == ndev
    assert after["direct_downloads"] == before["direct_downloads"]
    assert after["direct_uploads"] == before["direct_uploads"]
    single = KDTree(pts, leafsize=64, boxsize=1.0)
    d1, i1 = single.query(q[:200_000], k=8)
    assert np.array_equal(d[:200_000], d1) and np.array_equal(i[:200_000], i1)


def test_device_arrays_must_live_on_the_trees_device(gpu):
    torch = pytest.importorskip("torch")
    from nbodyhpc_b200.kdtree import KDTree

    pts = torch.rand((50_000, 3), device="cuda:0")
    tree = KDTree(pts, leafsize=64, boxsize=1.0)
    d, i = tree.query(pts[:100], k=2, return_squared=True)
    assert d.is_cuda and bool((d[:, 0] == 0).all())
    if gpu.lib().nbk_device_count() >= 2:
        with pytest.raises(RuntimeError, match="lives on GPU 1"):
            tree.query(pts[:100].to("cuda:1"), k=2)
        other = KDTree(pts.to("cuda:1"), leafsize=64, boxsize=1.0)  # follows the array's device
        assert other.device == 1


def test_python_result_buffers_are_recycled_pinned(gpu):
    """Result arrays a caller has dropped are kept (page-locked in the background) and back the next result
    of the same shape: that call then needs no staging copy, and the rows are the same."""
    import time

    from nbodyhpc_b200.kdtree import KDTree

    pts = philox(200_000, 42)
    q = np.random.default_rng(5).random((2_200_000, 3), dtype=np.float32)  # 2 x 70 MB of results at k = 8
    tree = KDTree(pts, leafsize=64, boxsize=1.0)
    s0 = gpu.host_path_stats()
    d, i = tree.query(q, k=8)
    s1 = gpu.host_path_stats()
    assert s1["staged_downloads"] == s0["staged_downloads"] + 1  # fresh pageable arrays: staged through the ring
    keep_d, keep_i = d.copy(), i.copy()
    released = {d.ctypes.data, i.ctypes.data}
    del d, i
    recycled_pinned = False
    for _ in range(8):  # the background thread page-locks the released buffers (milliseconds; be patient on a busy box)
        time.sleep(1.0)
        s1 = gpu.host_path_stats()
        d, i = tree.query(q, k=8)
        s2 = gpu.host_path_stats()
        assert np.array_equal(d, keep_d) and np.array_equal(i, keep_i)
        same_memory = {d.ctypes.data, i.ctypes.data} & released
        unstaged = s2["staged_downloads"] == s1["staged_downloads"] and s2["direct_downloads"] == s1["direct_downloads"]
        if same_memory and unstaged:
            recycled_pinned = True  # same memory, new arrays, written by the copy engine directly
            break
        released |= {d.ctypes.data, i.ctypes.data}
        del d, i
    assert recycled_pinned
    # results still referenced are never touched by later calls
    d2, i2 = tree.query(q, k=8)
    assert d2.ctypes.data != d.ctypes.data and np.array_equal(d2, d) and np.array_equal(d, keep_d)


@pytest.mark.parametrize("variant", [{"NBK_KERNEL": "packet"}, {"NBK_ORDER": "passes"}, {"NBK_MAX_SHARED_K": "8"},
                                     {"NBK_CDF_KEYS": "64"}, {"NBK_MORTON_FIRST_BIT": "0"}])
def test_runtime_selected_alternatives_agree(gpu, tmp_path, variant):
    """The alternatives kept behind environment switches (the packet kernel, the three-kernel ordering passes,
    the global-memory heap for k > 8, (d2, index) keys in the CDF's first pass, a 30-bit Morton order) must
    keep giving the rows and histograms of the defaults."""
    import os
    import subprocess
    import sys

    pts, q = philox(60_000, 42), philox(9_000, 43)
    np.save(tmp_path / "pts.npy", pts)
    np.save(tmp_path / "q.npy", q)
    script = (
        "import sys, numpy as np; sys.path.insert(0, %r)\n"
        "from nbodyhpc_b200 import capi\n"
        "pts, q = np.load(%r), np.load(%r)\n"
        "edges = np.linspace(0, 0.2, 25).astype(np.float32)\n"
        "out = {}\n"
        "for box in (None, 1.0):\n"
        "    t = capi.Tree.build(pts, 64, box)\n"
        "    for k in (1, 3, 8, 16):\n"
        "        d, i = t.query(q, k, squared=(k == 3))\n"
        "        out[f'd{k}{box}'] = d; out[f'i{k}{box}'] = i\n"
        "    if %r: out[f'cdf{box}'] = t.knn_cdf(q, [2, 8, 16], edges)\n"
        "np.savez(%r, **out)\n"
    )
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = {}
    for mode, extra in (("default", {}), ("variant", variant)):
        dst = str(tmp_path / f"{mode}.npz")
        env = {k: v for k, v in os.environ.items() if not k.startswith("NBK_")}
        env.update(extra)
        with_cdf = "NBK_KERNEL" not in variant  # the packet kernel has no CDF epilogue (it says so)
        subprocess.run([sys.executable, "-c",
                        script % (root, str(tmp_path / "pts.npy"), str(tmp_path / "q.npy"), with_cdf, dst)],
                       check=True, env=env, timeout=180)
        res[mode] = np.load(dst)
    assert len(res["default"].files) >= 16
    for key in res["default"].files:
        assert np.array_equal(res["default"][key], res["variant"][key]), (variant, key)
