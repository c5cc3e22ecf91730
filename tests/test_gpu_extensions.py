"""GPU tests of the rows SURVEY.md 8(f) marks "next": the fused kNN-CDF epilogue, device-array
ingest/egress (``__cuda_array_interface__``), in-process multi-GPU replicas and the CLI.

None of these exist in the reference; each is pinned to the drop-in path it extends:
  query_cdf(points, ks, bins)[i] == numpy.histogram(query(points, max(ks))[0][:, ks[i]-1], bins)[0]
  device arrays in  -> bit-identical results to host arrays in
  replicas          -> byte-identical rows to the single-GPU answer
"""
import os
import subprocess

import numpy as np
import pytest

from helpers import assert_parity, checker_tree, philox

pytestmark = pytest.mark.gpu


def _expected_cdf(dist, ks, edges):
    return np.stack([np.histogram(dist[:, k - 1], bins=edges)[0] for k in ks]).astype(np.uint64)


@pytest.mark.parametrize("box", [None, 1.0])
@pytest.mark.parametrize("ks", [[1, 2, 4, 8, 16, 32], [3, 5], [8], [64, 1], [7, 16]])
def test_knn_cdf_equals_histogram_of_rows(gpu, box, ks):
    n, m = 150_000, 40_000
    pts, q = philox(n, 42, box or 1.0), philox(m, 43, box or 1.0)
    tree = gpu.Tree.build(pts, 64, box)
    dist, _ = tree.query(q, max(ks))
    # edges that cut through the populated range, plus an empty bin on either side
    lo, hi = float(dist[:, min(ks) - 1].min()), float(dist[:, max(ks) - 1].max())
    edges = np.concatenate([[0.0], np.geomspace(max(lo, 1e-4) * 0.9, hi * 0.8, 37), [hi * 10]]).astype(np.float32)
    counts = tree.knn_cdf(q, ks, edges)
    assert counts.shape == (len(ks), len(edges) - 1)
    assert np.array_equal(counts, _expected_cdf(dist, ks, edges))
    # accumulates: a second call adds
    assert counts.sum() > 0


def test_knn_cdf_edge_semantics(gpu):
    """Values exactly on an edge go to the bin on their right, the last edge is inclusive, values
    outside are dropped -- numpy.histogram's convention."""
    pts = philox(20_000, 1)
    q = pts[:5000].copy()  # self-query: d(1st) == 0 exactly
    tree = gpu.Tree.build(pts, 32)
    dist, _ = tree.query(q, 2)
    d2 = np.sort(dist[:, 1])
    edges = np.array([0.0, d2[100], d2[2500], d2[-1]], np.float32)  # data values as edges, max as last edge
    counts = tree.knn_cdf(q, [1, 2], edges)
    assert np.array_equal(counts, _expected_cdf(dist, [1, 2], edges))
    assert counts[0, 0] == len(q) and counts[1].sum() == len(q)
    with pytest.raises(gpu.NbkError, match="distinct"):
        tree.knn_cdf(q, [2, 2], edges)
    with pytest.raises(gpu.NbkError, match="k must be positive integer"):
        tree.knn_cdf(q, [0], edges)
    with pytest.raises(gpu.NbkError, match="monotonically"):
        tree.knn_cdf(q, [1], np.array([0.0, 0.2, 0.1], np.float32))


def test_python_query_cdf(gpu):
    from nbodyhpc_b200.kdtree import KDTree

    pts, q = philox(100_000, 4), philox(30_000, 5)
    tree = KDTree(pts, leafsize=64, boxsize=1.0)
    ks = [1, 4, 8]
    edges = np.linspace(0.0, 0.08, 33).astype(np.float32)
    counts = tree.query_cdf(q, ks, edges)
    dist, _ = tree.query(q, k=8)
    assert np.array_equal(counts, _expected_cdf(dist, ks, edges))


# ---- device arrays -----------------------------------------------------------------------------------
class _RawDeviceView:
    """A minimal non-torch ``__cuda_array_interface__`` provider (what cupy / numba arrays look like)."""

    def __init__(self, tensor):
        self._keep = tensor
        self.__cuda_array_interface__ = {"shape": tuple(tensor.shape), "typestr": "<f4",
                                         "data": (tensor.data_ptr(), False), "version": 3, "strides": None}


def test_device_arrays_torch_roundtrip(gpu):
    import torch

    from nbodyhpc_b200.kdtree import KDTree

    pts, q = philox(200_000, 42), philox(50_000, 43)
    host_tree = KDTree(pts, leafsize=64, boxsize=1.0)
    d_host, i_host = host_tree.query(q, k=8)

    t_pts, t_q = torch.from_numpy(pts).cuda(), torch.from_numpy(q).cuda()
    dev_tree = KDTree(t_pts, leafsize=64, boxsize=1.0)
    assert dev_tree.n == host_tree.n and dev_tree.size == host_tree.size
    assert np.array_equal(dev_tree.nodes(), host_tree.nodes())
    d_dev, i_dev = dev_tree.query(t_q, k=8)
    assert d_dev.is_cuda and d_dev.shape == (len(q), 8) and i_dev.dtype == torch.int32
    torch.cuda.synchronize()
    assert np.array_equal(d_dev.cpu().numpy(), d_host)
    assert np.array_equal(i_dev.cpu().numpy().view(np.uint32), i_host)
    # N-d queries keep their leading shape
    d3, i3 = dev_tree.query(t_q.reshape(100, 500, 3), k=2)
    assert tuple(d3.shape) == (100, 500, 2)
    assert np.array_equal(d3.cpu().numpy().reshape(-1, 2), d_host[:, :2])
    # the CDF from device-resident queries
    edges = np.linspace(0.0, 0.05, 17).astype(np.float32)
    assert np.array_equal(dev_tree.query_cdf(t_q, [1, 8], edges), _expected_cdf(d_host, [1, 8], edges))
    # refusals instead of silent host copies
    with pytest.raises(TypeError, match="float32"):
        dev_tree.query(t_q.double(), k=1)
    with pytest.raises(TypeError, match="contiguous"):
        dev_tree.query(t_q[::2], k=1)


def test_device_arrays_generic_interface(gpu):
    import torch

    from nbodyhpc_b200.kdtree import DeviceArray, KDTree

    pts, q = philox(50_000, 7), philox(5_000, 8)
    tree = KDTree(_RawDeviceView(torch.from_numpy(pts).cuda()), leafsize=32)
    d, i = tree.query(_RawDeviceView(torch.from_numpy(q).cuda()), k=4)
    assert isinstance(d, DeviceArray) and d.shape == (len(q), 4)
    d_ref, i_ref = checker_tree(pts, 32, None).query(q, 4, workers=0)
    assert_parity(d.to_host(), i.to_host(), d_ref, i_ref, pts, q, None)
    # torch can adopt the result without a copy
    t = torch.as_tensor(d, device="cuda")
    assert np.array_equal(t.cpu().numpy(), d.to_host())


def test_nd_query_reshape_fix(gpu):
    """The reference's N-d reshape raises TypeError (__init__.py:52-54); the drop-in returns (..., k)."""
    from nbodyhpc_b200.kdtree import KDTree

    pts = philox(10_000, 3)
    tree = KDTree(pts)
    q = philox(600, 9).reshape(20, 30, 3)
    d, i = tree.query(q, k=3)
    assert d.shape == (20, 30, 3) and i.shape == (20, 30, 3)
    d2, i2 = tree.query(q.reshape(-1, 3), k=3)
    assert np.array_equal(d.reshape(-1, 3), d2) and np.array_equal(i.reshape(-1, 3), i2)


# ---- CLI ------------------------------------------------------------------------------------------------
def test_kdtree_main_cli(gpu, tmp_path):
    """The reference's CLI report (main.cpp:161-174) over the drop-in: self-query, nearest = self."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "nbodyhpc_b200", "lib", "kdtree_main")
    for extra in ([], ["--periodic", "--box_size", "1.0"]):
        out = subprocess.run([exe, "-n", "200000", "-q", "20000", "--num-neighbors", "8", "--leaf-size", "64", *extra],
                             capture_output=True, text=True, timeout=120)
        assert out.returncode == 0, out.stderr
        lines = out.stdout.strip().splitlines()
        assert lines[0] == "Benchmarking kdtree with 200000 points"
        assert "Total distance was not 0" not in out.stdout
        keys = [l.split(":")[0] for l in lines[1:]]
        assert keys == ["Build time", "Query time", "Query performance", "Points visited proportion"]
        visited = float(lines[-1].split(":")[1].strip().rstrip("%"))
        # the reference visits ~256 points per k=8 query at leaf 64: ~0.13 % of 200k
        assert 0.05 < visited < 0.4
    # file input: raw float32 xyz
    pts = philox(50_000, 3)
    path = str(tmp_path / "cli_points.f32")
    pts.tofile(path)
    out = subprocess.run([exe, "-f", path, "-q", "1000", "--num-neighbors", "4"], capture_output=True, text=True,
                         timeout=120)
    assert out.returncode == 0 and out.stdout.startswith("Benchmarking kdtree with data from:"), out.stderr
    bad = subprocess.run([exe, "--no-such-flag"], capture_output=True, text=True, timeout=60)
    assert bad.returncode == 1 and "unknown option" in bad.stderr


# ---- host pipeline ----------------------------------------------------------------------------------------
def test_staged_download_into_pageable_memory_matches_pinned(gpu):
    """Large results for ordinary numpy arrays go through the pinned ring + host threads
    (host_stage.cuh); pinned destinations are written by the copy engine directly.  Same bytes."""
    import torch

    pts, q = philox(100_000, 42), philox(2_300_000, 43)
    tree = gpu.Tree.build(pts, 64, 1.0)
    k = 8
    d_page, i_page = tree.query(q, k)  # 147 MB of results: staged
    q_pin = torch.from_numpy(q).pin_memory()
    d_pin = torch.empty((len(q), k), dtype=torch.float32).pin_memory()
    i_pin = torch.empty((len(q), k), dtype=torch.int32).pin_memory()
    tree.query_raw(q_pin.data_ptr(), len(q), k, d_pin.data_ptr(), i_pin.data_ptr())
    assert np.array_equal(d_page, d_pin.numpy()) and np.array_equal(i_page, i_pin.numpy().view(np.uint32))
    # and through the Python drop-in (library-allocated, huge-page advised result arrays)
    from nbodyhpc_b200.kdtree import KDTree

    d_py, i_py = KDTree(pts, leafsize=64, boxsize=1.0).query(q, k=k)
    assert d_py.flags["C_CONTIGUOUS"] and d_py.flags["OWNDATA"] is False and d_py.base is not None
    assert np.array_equal(d_py, d_page) and np.array_equal(i_py, i_page)


# ---- the C++ drop-in, exercised like the reference's own gtest suite ------------------------------------------
def test_cpp_dropin_replays_reference_cpp_tests(gpu, tmp_path):
    """tests/cpp/test_kdtree_dropin.cpp = kdtree/src/cpp/tests/test.cpp:43-114 against
    include/kdtree/*.hpp (wenda::kdtree::KDTree over the C ABI): built and run here."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "nbodyhpc_b200", "lib")
    exe = str(tmp_path / "test_kdtree_dropin")
    subprocess.run(["g++", "-O2", "-std=c++20", "-I", os.path.join(root, "include"),
                    os.path.join(root, "tests", "cpp", "test_kdtree_dropin.cpp"), "-o", exe, "-L", libdir, "-lnbk",
                    f"-Wl,-rpath,{libdir}"], check=True, timeout=300)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "all checks passed" in out.stdout, out.stdout + out.stderr


def test_concurrent_queries_on_one_tree(gpu):
    """A const KDTree is safe for concurrent find_closest in the reference (kdtree.hpp:207-210, the pybind
    thread pool relies on it); the same must hold for concurrent batched queries on one device tree."""
    from concurrent.futures import ThreadPoolExecutor

    pts = philox(300_000, 42)
    tree = gpu.Tree.build(pts, 64, 1.0)
    batches = [philox(400_000 + 1000 * j, 50 + j) for j in range(6)]
    ks = [1, 8, 16, 8, 3, 32]
    serial = [tree.query(q, k) for q, k in zip(batches, ks)]
    with ThreadPoolExecutor(max_workers=6) as pool:
        parallel = list(pool.map(lambda a: tree.query(a[0], a[1]), zip(batches, ks)))
    for (d0, i0), (d1, i1) in zip(serial, parallel):
        assert np.array_equal(d0, d1) and np.array_equal(i0, i1)
    # builds may run concurrently with queries too
    with ThreadPoolExecutor(max_workers=3) as pool:
        f_build = pool.submit(lambda: gpu.Tree.build(philox(200_000, 7), 32))
        f_query = pool.submit(lambda: tree.query(batches[1], 8))
        other, (d, i) = f_build.result(), f_query.result()
    assert np.array_equal(d, serial[1][0]) and other.size > 1
