"""CPU-side checks of the boundary: libnbk.so loads and exports every symbol include/nbk.h
declares, its host-only logic (topology plan, argument validation, error wording) matches the
reference, and compute entry points fail loudly -- never fall back -- without a GPU."""
import os
import re
import subprocess

import numpy as np
import pytest

from helpers import Oracle, Reference, philox

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "nbk.h")).read()
    return sorted(set(re.findall(r"NBK_API[^;(]*?\b(nbk_\w+)\s*\(", text)))


def test_header_symbols_are_exported(native):
    declared = declared_symbols()
    assert sorted(native.SYMBOLS) == declared
    out = subprocess.run(["nm", "-D", "--defined-only", native.library_path()], capture_output=True, text=True,
                         check=True).stdout
    exported = sorted(set(re.findall(r"\bT (nbk_\w+)", out)))
    assert exported == declared
    lib = native.lib()
    for name in declared:
        assert hasattr(lib, name)


def test_library_contains_sm100a_code_only(native):
    out = subprocess.run(["cuobjdump", "-lelf", native.library_path()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_node_and_meta_layout(native):
    import ctypes as C

    assert native.NODE_DTYPE.itemsize == 16  # KDTreeNode, kdtree.hpp:149-163
    assert C.sizeof(native.TreeMeta) == 80


@pytest.mark.parametrize("n", [0, 1, 8, 10, 16, 17, 100, 1000, 4097, 100_003, 1_000_000, 2 ** 21])
@pytest.mark.parametrize("leaf", [1, 16, 32, 64, 128])
def test_topology_plan_matches_reference_rule(native, n, leaf):
    nodes, n_nodes, n_levels = native.plan_topology(n, leaf)
    assert n_nodes == Oracle.expected_num_nodes(n, leaf)
    leaves = nodes[nodes["dim"] == -1]
    sizes = leaves["right"].astype(np.int64) - leaves["left"]
    n_pad = (n + 7) // 8 * 8
    assert sizes.sum() == n_pad and (sizes % 8 == 0).all() and (sizes <= max(leaf, 16)).all()
    # leaves tile [0, n_pad) in pre-order
    assert np.array_equal(leaves["left"][1:], leaves["right"][:-1])
    if n <= 100_003:
        ref_nodes = Oracle.Tree(philox(n, 1), leaf, None).nodes()
        for f in ("dim", "left", "right"):
            assert np.array_equal(nodes[f], ref_nodes[f])


def test_headline_topology(native):
    # 512^3 / leaf 64 -> 4 194 303 nodes, 21 split levels; leaf 128 -> 2 097 151
    assert native.plan_topology(512 ** 3, 64, with_nodes=False)[1:] == (4194303, 21)
    assert native.plan_topology(512 ** 3, 128, with_nodes=False)[1:] == (2097151, 20)
    assert native.plan_topology(1_000_000, 64, with_nodes=False)[1] == 32767


def test_argument_errors_use_reference_wording(native):
    with pytest.raises(native.NbkError, match="block_size must be a multiple of 8."):
        native.plan_topology(100, 64, block_size=12)
    with pytest.raises(native.NbkError, match="More than uint32_t points are not supported."):
        native.plan_topology(2 ** 32 + 5, 64)
    with pytest.raises(native.NbkError, match=r"positions must be a 2D array of shape \(N, 3\)"):
        native.Tree.build(np.zeros((4, 2), np.float32))


def test_no_cpu_fallback(native):
    if native.lib().nbk_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(native.NbkError) as err:
        native.Tree.build(philox(100, 1), 64)
    assert err.value.code == native.NBK_ERR_CUDA and "no CPU fallback" in str(err.value)
    from nbodyhpc_b200.kdtree import KDTree

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        KDTree(philox(100, 1))


def test_python_layer_does_not_import_the_oracle():
    """The product path must not route through oracle/ (or any CPU implementation)."""
    for root, _, files in os.walk(os.path.join(ROOT, "nbodyhpc_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(root, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "liboracle" not in text and "libnbref" not in text, f


def test_shard_range_is_the_thread_pool_split():
    from nbodyhpc_b200.dist import shard_range

    for total in (0, 1, 7, 8, 100, 10 ** 8 + 3):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and max(e for _, e in spans) == total
            covered = sum(e - b for b, e in spans)
            assert covered == total
            for (b0, e0), (b1, e1) in zip(spans, spans[1:]):
                assert e0 == b1 or (e0 - b0 == 0 or e1 - b1 == 0) or e0 <= b1
            if total >= world:
                assert all(e - b == total // world for b, e in spans[:-1])


def test_cpp_fixture_generators_match_the_reference(native):
    """include/kdtree/kdtree_utils.hpp restates Random123's Philox4x32-10 (not vendored here): its
    make_random_position_and_index must give the values of the reference's (kdtree_utils.hpp:16-46),
    which the oracle's generator is pinned to (tests/test_oracle.py, golden vectors)."""
    from helpers import Oracle
    from nbodyhpc_b200.kdtree import _impl

    for n, seed, box in [(10, 42, 1.0), (1000, 43, 2.0), (4097, 7, 1.0)]:
        assert np.array_equal(_impl._make_random_positions(n, seed, box), Oracle.philox_points(n, seed, box))
    # the CLI's layout (main.cpp:14-35): counter {i}, lanes 0..2 -- in (0, 1], deterministic
    a, b = _impl._fill_random_positions(1000, 42), _impl._fill_random_positions(1000, 42)
    assert np.array_equal(a, b) and a.min() > 0.0 and a.max() <= 1.0 and len(np.unique(a)) > 2990


def test_cpp_containers_and_helpers_host_only(native, tmp_path):
    """tests/cpp/test_containers.cpp: PositionAndIndexArray proxy/random-access iterators under STL and
    ranges algorithms, OffsetRangeContainerWrapper, make_position_and_indices, <span.hpp> -- the parts of
    the reference's public headers (position_array.hpp:26-161,273-352; kdtree_utils.hpp:117-118;
    kdtree.hpp:11) that need no device."""
    libdir = os.path.join(ROOT, "nbodyhpc_b200", "lib")
    exe = str(tmp_path / "test_containers")
    subprocess.run(["g++", "-O1", "-std=c++20", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "test_containers.cpp"), "-o", exe, "-L", libdir, "-lnbk",
                    f"-Wl,-rpath,{libdir}"], check=True, timeout=300)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and "all checks passed" in out.stdout, out.stdout + out.stderr


def test_query_flag_and_k_validation_without_device(native):
    """Argument errors come before any device work and use the reference's wording."""
    q = np.zeros((3, 3), np.float32)
    with pytest.raises(native.NbkError, match="k must be positive integer"):
        native.scan_block(q[:0, 0], q[:0, 1], q[:0, 2], np.zeros(0, np.uint32), q, 0)
    with pytest.raises(native.NbkError, match="block_size must be a multiple of 8."):
        native.scan_block(q[:, 0], q[:, 1], q[:, 2], np.zeros(3, np.uint32), q, 1)
    stats = native.host_path_stats()
    assert set(stats) == {"staged_downloads", "direct_downloads", "staged_uploads", "direct_uploads"}
    assert native.pointer_device(q.ctypes.data) == -1
