"""ctypes view of the C ABI (include/nbk.h -> lib/libnbk.so).

This is the same boundary a cgo / JNI / ctypes binding in the reference would use.  It exists for
the callers that work with raw device pointers (bench.py, nbodyhpc_b200.dist, the GPU parity tests);
the numpy-facing API is ``nbodyhpc_b200.kdtree.KDTree`` (pybind11, same library underneath).
torch appears nowhere in here: device buffers cross as integer addresses.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

NBK_OK, NBK_ERR_INVALID, NBK_ERR_CUDA, NBK_ERR_NOMEM = 0, 1, 2, 3

NODE_DTYPE = np.dtype([("dim", np.int32), ("split", np.float32), ("left", np.uint32), ("right", np.uint32)])

# every symbol include/nbk.h declares (tests check that the library exports exactly these)
SYMBOLS = [
    "nbk_last_error", "nbk_launch_count", "nbk_device_count",
    "nbk_tree_build", "nbk_tree_build_device", "nbk_tree_build_soa", "nbk_plan_topology", "nbk_tree_free",
    "nbk_tree_get_meta", "nbk_tree_device", "nbk_tree_copy_nodes", "nbk_tree_copy_points",
    "nbk_tree_query", "nbk_tree_query_device", "nbk_tree_query_ex", "nbk_tree_query_ex2",
    "nbk_tree_query_device_ex", "nbk_scan_block", "nbk_tree_stats",
    "nbk_tree_knn_cdf", "nbk_tree_knn_cdf_device",
    "nbk_tree_arena", "nbk_tree_alloc_replica", "nbk_tree_clone_to_device", "nbk_profile_enable", "nbk_profile_read",
    "nbk_device_alloc", "nbk_device_alloc_on", "nbk_pointer_device", "nbk_device_free", "nbk_device_copy",
    "nbk_device_zero", "nbk_host_path_stats", "nbk_host_alloc", "nbk_host_free", "nbk_host_register",
    "nbk_host_unregister",
]

SECTION_QUERY_ORDER, SECTION_KNN_KERNEL = 0, 1
QUERY_SQUARED = 1  # NBK_QUERY_SQUARED


class TreeMeta(C.Structure):
    _fields_ = [
        ("n_points", C.c_uint64), ("n_padded", C.c_uint64), ("n_nodes", C.c_uint64),
        ("arena_bytes", C.c_uint64), ("leaf_size", C.c_int32), ("block_size", C.c_int32),
        ("periodic", C.c_int32), ("box_size", C.c_float), ("lo", C.c_float * 3), ("hi", C.c_float * 3),
        ("n_levels", C.c_int32), ("reserved", C.c_int32),
    ]


def library_path() -> str:
    # NBK_LIBRARY: an alternative build of the same library (kernel-variant experiments)
    override = os.environ.get("NBK_LIBRARY")
    if override:
        return override
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libnbk.so")


_lib = None


def lib() -> C.CDLL:
    """Loads libnbk.so; raises (never falls back) if it has not been built."""
    global _lib
    if _lib is None:
        path = library_path()
        if not os.path.exists(path):
            raise ImportError(f"{path} is not built; run `python -m nbodyhpc_b200._build`. There is no CPU fallback.")
        L = C.CDLL(path)
        vp, u64, i32, f32 = C.c_void_p, C.c_uint64, C.c_int, C.c_float
        ip = C.POINTER(C.c_int)
        L.nbk_last_error.restype = C.c_char_p
        L.nbk_launch_count.restype = u64
        L.nbk_device_count.restype = i32
        L.nbk_tree_build.restype = vp
        L.nbk_tree_build.argtypes = [vp, u64, i32, i32, i32, f32, i32, ip]
        L.nbk_tree_build_device.restype = vp
        L.nbk_tree_build_device.argtypes = [vp, u64, i32, i32, i32, f32, i32, vp, ip]
        L.nbk_tree_build_soa.restype = vp
        L.nbk_tree_build_soa.argtypes = [vp, vp, vp, vp, u64, i32, i32, i32, f32, i32, ip]
        L.nbk_plan_topology.argtypes = [u64, i32, i32, vp, C.POINTER(u64), ip]
        L.nbk_tree_free.argtypes = [vp]
        L.nbk_tree_get_meta.argtypes = [vp, C.POINTER(TreeMeta)]
        L.nbk_tree_device.argtypes = [vp]
        L.nbk_tree_copy_nodes.argtypes = [vp, vp]
        L.nbk_tree_copy_points.argtypes = [vp, vp, vp, vp, vp]
        L.nbk_tree_query.argtypes = [vp, vp, u64, i32, vp, vp]
        L.nbk_tree_query_device.argtypes = [vp, vp, u64, i32, vp, vp, vp]
        L.nbk_tree_query_ex.argtypes = [vp, vp, u64, i32, i32, f32, vp, vp]
        L.nbk_tree_query_ex2.argtypes = [vp, vp, u64, i32, i32, f32, i32, vp, vp]
        L.nbk_tree_query_device_ex.argtypes = [vp, vp, u64, i32, i32, f32, i32, vp, vp, vp]
        L.nbk_scan_block.argtypes = [vp, vp, vp, vp, u64, vp, u64, i32, i32, f32, i32, vp, vp, i32]
        L.nbk_tree_stats.argtypes = [vp, vp, u64, i32, i32, f32, vp]
        L.nbk_tree_knn_cdf.argtypes = [vp, vp, u64, vp, i32, vp, i32, vp]
        L.nbk_tree_knn_cdf_device.argtypes = [vp, vp, u64, vp, i32, vp, i32, vp, vp]
        L.nbk_tree_arena.argtypes = [vp, C.POINTER(vp), C.POINTER(u64)]
        L.nbk_tree_alloc_replica.restype = vp
        L.nbk_tree_alloc_replica.argtypes = [C.POINTER(TreeMeta), i32, ip]
        L.nbk_profile_enable.argtypes = [i32]
        L.nbk_profile_read.argtypes = [i32, C.POINTER(C.c_double), C.POINTER(u64)]
        L.nbk_tree_clone_to_device.restype = vp
        L.nbk_tree_clone_to_device.argtypes = [vp, i32, ip]
        L.nbk_device_alloc.restype = vp
        L.nbk_device_alloc.argtypes = [u64]
        L.nbk_device_alloc_on.restype = vp
        L.nbk_device_alloc_on.argtypes = [i32, u64]
        L.nbk_pointer_device.argtypes = [vp]
        L.nbk_host_path_stats.argtypes = [vp]
        L.nbk_device_free.argtypes = [vp]
        L.nbk_device_copy.argtypes = [vp, vp, u64, i32]
        L.nbk_device_zero.argtypes = [vp, u64]
        L.nbk_host_alloc.restype = vp
        L.nbk_host_alloc.argtypes = [u64]
        L.nbk_host_free.argtypes = [vp]
        L.nbk_host_register.argtypes = [vp, u64, i32]
        L.nbk_host_unregister.argtypes = [vp]
        _lib = L
    return _lib


class NbkError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(message)
        self.code = code


def _check(status: int) -> None:
    if status != NBK_OK:
        raise NbkError(status, lib().nbk_last_error().decode())


def launch_count() -> int:
    return int(lib().nbk_launch_count())


def plan_topology(n_points: int, leaf_size: int = 64, block_size: int = 8, with_nodes: bool = True):
    """Host-only: (node records with split = 0, number of levels) the build will produce."""
    n_nodes, n_levels = C.c_uint64(), C.c_int()
    _check(lib().nbk_plan_topology(n_points, leaf_size, block_size, None, C.byref(n_nodes), C.byref(n_levels)))
    nodes = None
    if with_nodes:
        nodes = np.empty(n_nodes.value, NODE_DTYPE)
        _check(lib().nbk_plan_topology(n_points, leaf_size, block_size, _host_ptr(nodes), None, None))
    return nodes, int(n_nodes.value), int(n_levels.value)


def device_alloc(nbytes: int, device: int = -1) -> int:
    """cudaMalloc on ``device`` (-1: the calling thread's current device)."""
    p = lib().nbk_device_alloc_on(device, nbytes)
    if not p:
        raise NbkError(NBK_ERR_NOMEM, lib().nbk_last_error().decode())
    return int(p)


def pointer_device(ptr: int) -> int:
    """Device ordinal owning ``ptr``; -1 for host or unknown memory."""
    return int(lib().nbk_pointer_device(C.c_void_p(ptr)))


def host_path_stats() -> dict:
    """How large pageable host buffers crossed the boundary so far (staged through pinned memory or not)."""
    out = np.zeros(4, np.uint64)
    _check(lib().nbk_host_path_stats(_host_ptr(out)))
    return dict(zip(("staged_downloads", "direct_downloads", "staged_uploads", "direct_uploads"), map(int, out)))


def scan_block(x, y, z, idx, q, k: int, boxsize=None, squared: bool = False, device: int = -1):
    """The leaf scan + top-k container on ONE flat block of points (no tree): nbk_scan_block."""
    x, y, z = (np.ascontiguousarray(a, dtype=np.float32) for a in (x, y, z))
    idx = np.ascontiguousarray(idx, dtype=np.uint32)
    q = np.ascontiguousarray(q, dtype=np.float32)
    m = q.shape[0]
    d = np.empty((m, max(k, 0)), np.float32)
    i = np.empty((m, max(k, 0)), np.uint32)
    _check(lib().nbk_scan_block(_host_ptr(x), _host_ptr(y), _host_ptr(z), _host_ptr(idx), x.shape[0], _host_ptr(q), m, k,
                                int(boxsize is not None), float(boxsize or 0.0), QUERY_SQUARED if squared else 0,
                                _host_ptr(d), _host_ptr(i), device))
    return d, i


def device_free(ptr: int) -> None:
    lib().nbk_device_free(C.c_void_p(ptr))


def host_to_device(d_ptr: int, a: np.ndarray) -> None:
    a = np.ascontiguousarray(a)
    _check(lib().nbk_device_copy(C.c_void_p(d_ptr), _host_ptr(a), a.nbytes, 0))


def device_to_host(out: np.ndarray, d_ptr: int) -> None:
    _check(lib().nbk_device_copy(_host_ptr(out), C.c_void_p(d_ptr), out.nbytes, 1))


def device_memset(d_ptr: int, nbytes: int) -> None:
    _check(lib().nbk_device_zero(C.c_void_p(d_ptr), nbytes))


def profile_enable(on: bool) -> None:
    lib().nbk_profile_enable(int(on))


def profile_read(section: int):
    """(total milliseconds, recordings) of a section since the last read; waits for the events."""
    ms, cnt = C.c_double(), C.c_uint64()
    _check(lib().nbk_profile_read(section, C.byref(ms), C.byref(cnt)))
    return ms.value, int(cnt.value)


def _host_ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class Tree:
    """Owns one ``nbk_tree*``.  Host arrays are numpy; device buffers are integer addresses."""

    def __init__(self, handle: int, owned: bool = True):
        if not handle:
            raise ValueError("null tree handle")
        self._h = C.c_void_p(handle)
        self._owned = owned  # False: a view of a tree owned elsewhere (the pybind object)

    # ---- construction ---------------------------------------------------------------------
    @classmethod
    def build(cls, points: np.ndarray, leaf_size: int = 64, boxsize=None, device: int = -1, block_size: int = 8):
        pts = np.ascontiguousarray(points, dtype=np.float32)
        if pts.ndim != 2 or pts.shape[1] != 3:
            raise NbkError(NBK_ERR_INVALID, "positions must be a 2D array of shape (N, 3)")
        status = C.c_int(0)
        h = lib().nbk_tree_build(_host_ptr(pts), pts.shape[0], leaf_size, block_size, int(boxsize is not None),
                                 float(boxsize or 0.0), device, C.byref(status))
        _check(status.value)
        return cls(h)

    @classmethod
    def build_device(cls, d_points: int, n: int, leaf_size: int = 64, boxsize=None, device: int = -1,
                     stream: int = 0, block_size: int = 8):
        status = C.c_int(0)
        h = lib().nbk_tree_build_device(C.c_void_p(d_points), n, leaf_size, block_size, int(boxsize is not None),
                                        float(boxsize or 0.0), device, C.c_void_p(stream), C.byref(status))
        _check(status.value)
        return cls(h)

    @classmethod
    def build_soa(cls, x, y, z, idx, leaf_size: int = 64, boxsize=None, device: int = -1, block_size: int = 8):
        x, y, z = (np.ascontiguousarray(a, dtype=np.float32) for a in (x, y, z))
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        status = C.c_int(0)
        h = lib().nbk_tree_build_soa(_host_ptr(x), _host_ptr(y), _host_ptr(z), _host_ptr(idx), x.shape[0], leaf_size,
                                     block_size, int(boxsize is not None), float(boxsize or 0.0), device,
                                     C.byref(status))
        _check(status.value)
        return cls(h)

    @classmethod
    def alloc_replica(cls, meta: TreeMeta, device: int = -1):
        status = C.c_int(0)
        h = lib().nbk_tree_alloc_replica(C.byref(meta), device, C.byref(status))
        _check(status.value)
        return cls(h)

    def clone_to_device(self, device: int):
        """A byte-identical replica on another GPU of this process (peer copy)."""
        status = C.c_int(0)
        h = lib().nbk_tree_clone_to_device(self._h, device, C.byref(status))
        _check(status.value)
        return type(self)(h)

    def close(self):
        if getattr(self, "_h", None):
            if getattr(self, "_owned", True):
                lib().nbk_tree_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- introspection --------------------------------------------------------------------
    @property
    def meta(self) -> TreeMeta:
        m = TreeMeta()
        _check(lib().nbk_tree_get_meta(self._h, C.byref(m)))
        return m

    @property
    def n(self) -> int:
        return int(self.meta.n_padded)

    @property
    def size(self) -> int:
        return int(self.meta.n_nodes)

    @property
    def device(self) -> int:
        return int(lib().nbk_tree_device(self._h))

    def nodes(self) -> np.ndarray:
        out = np.empty(self.size, NODE_DTYPE)
        _check(lib().nbk_tree_copy_nodes(self._h, _host_ptr(out)))
        return out

    def points(self):
        n = self.n
        x, y, z = (np.empty(n, np.float32) for _ in range(3))
        idx = np.empty(n, np.uint32)
        _check(lib().nbk_tree_copy_points(self._h, _host_ptr(x), _host_ptr(y), _host_ptr(z), _host_ptr(idx)))
        return x, y, z, idx

    def arena(self):
        ptr, nbytes = C.c_void_p(), C.c_uint64()
        _check(lib().nbk_tree_arena(self._h, C.byref(ptr), C.byref(nbytes)))
        return int(ptr.value), int(nbytes.value)

    # ---- query ----------------------------------------------------------------------------
    def query(self, q: np.ndarray, k: int = 1, periodic: int = -1, boxsize: float = 0.0, out=None,
              squared: bool = False):
        q = np.ascontiguousarray(q, dtype=np.float32)
        if q.ndim != 2 or q.shape[1] != 3:
            raise NbkError(NBK_ERR_INVALID, "positions must be a 2D array of shape (N, 3)")
        m = q.shape[0]
        if out is None:
            d = np.empty((m, max(k, 0)), np.float32)
            i = np.empty((m, max(k, 0)), np.uint32)
        else:
            d, i = out
        _check(lib().nbk_tree_query_ex2(self._h, _host_ptr(q), m, k, periodic, boxsize, QUERY_SQUARED if squared else 0,
                                        _host_ptr(d), _host_ptr(i)))
        return d, i

    def query_raw(self, q_ptr: int, m: int, k: int, out_d_ptr: int, out_i_ptr: int):
        """Host-pointer entry point (nbk_tree_query) on raw addresses, e.g. pinned torch tensors."""
        _check(lib().nbk_tree_query(self._h, C.c_void_p(q_ptr), m, k, C.c_void_p(out_d_ptr), C.c_void_p(out_i_ptr)))

    def query_device(self, d_q: int, m: int, k: int, d_out_d: int, d_out_i: int, stream: int = 0,
                     periodic: int = -1, boxsize: float = 0.0, squared: bool = False):
        """Enqueues the query on ``stream`` (a cudaStream_t as int); does not synchronise."""
        _check(lib().nbk_tree_query_device_ex(self._h, C.c_void_p(d_q), m, k, periodic, boxsize,
                                              QUERY_SQUARED if squared else 0, C.c_void_p(d_out_d),
                                              C.c_void_p(d_out_i), C.c_void_p(stream)))

    def knn_cdf(self, q: np.ndarray, ks, edges) -> np.ndarray:
        """counts[i, b] = numpy.histogram(dist[:, ks[i]-1], edges)[0][b] without materialising the rows."""
        q = np.ascontiguousarray(q, dtype=np.float32)
        ks = np.ascontiguousarray(ks, dtype=np.int32)
        edges = np.ascontiguousarray(edges, dtype=np.float32)
        counts = np.zeros((len(ks), len(edges) - 1), np.uint64)
        _check(lib().nbk_tree_knn_cdf(self._h, _host_ptr(q), q.shape[0], _host_ptr(ks), len(ks), _host_ptr(edges),
                                      len(edges) - 1, _host_ptr(counts)))
        return counts

    def knn_cdf_device(self, d_q: int, m: int, ks, d_edges: int, n_bins: int, d_counts: int, stream: int = 0):
        ks = np.ascontiguousarray(ks, dtype=np.int32)
        _check(lib().nbk_tree_knn_cdf_device(self._h, C.c_void_p(d_q), m, _host_ptr(ks), len(ks), C.c_void_p(d_edges),
                                             n_bins, C.c_void_p(d_counts), C.c_void_p(stream)))

    def stats(self, q: np.ndarray, k: int = 1, periodic: int = -1, boxsize: float = 0.0) -> np.ndarray:
        q = np.ascontiguousarray(q, dtype=np.float32)
        out = np.zeros(3, np.uint64)
        _check(lib().nbk_tree_stats(self._h, _host_ptr(q), q.shape[0], k, periodic, boxsize, _host_ptr(out)))
        return out
