"""TEST INFRASTRUCTURE -- not part of the product path.

ctypes loaders for the two CPU checkers and the row-wise parity checker.

* ``Oracle``    -- oracle/liboracle.so, the plain-C restatement (oracle/knn_oracle.c).
* ``Reference`` -- oracle/_ref/libnbref.so, the reference's own sources compiled unmodified
  (oracle/Makefile, oracle/ref_harness.cpp).  ``Reference.available()`` is False when neither
  /root/reference nor a prebuilt _ref/ is present.
* ``compare_knn`` -- the parity contract of SURVEY.md section 8(c): distances bit-equal row by row,
  indices equal except where an exact distance tie makes the reference's answer order- or
  traversal-dependent, in which case the tie is verified by recomputation.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline / ``--impl reference`` legs
may import this package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")

NODE_DTYPE = np.dtype([("dim", np.int32), ("split", np.float32), ("left", np.uint32), ("right", np.uint32)])


def build_checkers(quiet: bool = True) -> None:
    """Runs oracle/Makefile (compiles the C restatement; and _ref when /root/reference exists)."""
    subprocess.run(["make", "-C", _HERE] + (["-s"] if quiet else []), check=True)


def _aos(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim != 2 or a.shape[1] != 3:
        raise ValueError("expected (N, 3) array")
    return a


class _TreeBase:
    """Common surface of the two checkers' trees."""

    lib = None
    prefix = ""

    def __init__(self, handle, periodic: bool, box: float):
        self._h = handle
        self.periodic = periodic
        self.boxsize = box

    def _fn(self, name):
        return getattr(self.lib, self.prefix + name)

    @property
    def n(self) -> int:
        return int(self._fn("tree_num_points")(self._h))

    @property
    def size(self) -> int:
        return int(self._fn("tree_num_nodes")(self._h))

    def nodes(self) -> np.ndarray:
        out = np.empty(self.size, dtype=NODE_DTYPE)
        self._fn("tree_copy_nodes")(self._h, out.ctypes.data_as(C.c_void_p))
        return out

    def points(self):
        n = self.n
        x, y, z = (np.empty(n, np.float32) for _ in range(3))
        idx = np.empty(n, np.uint32)
        self._fn("tree_copy_points")(self._h, x, y, z, idx)
        return x, y, z, idx

    def __del__(self):
        if getattr(self, "_h", None):
            self._fn("tree_free")(self._h)
            self._h = None


class Oracle:
    """The plain-C restatement."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            path = os.path.join(_HERE, "liboracle.so")
            if not os.path.exists(path):
                build_checkers()
            L = C.CDLL(path)
            L.orc_philox_points.argtypes = [C.c_uint32, C.c_uint32, C.c_float, _f32p]
            L.orc_tree_build.restype = C.c_void_p
            L.orc_tree_build.argtypes = [_f32p, C.c_uint64, C.c_int, C.c_int, C.c_float, C.POINTER(C.c_int)]
            L.orc_tree_free.argtypes = [C.c_void_p]
            L.orc_tree_num_points.restype = C.c_uint64
            L.orc_tree_num_points.argtypes = [C.c_void_p]
            L.orc_tree_num_nodes.restype = C.c_uint64
            L.orc_tree_num_nodes.argtypes = [C.c_void_p]
            L.orc_tree_copy_nodes.argtypes = [C.c_void_p, C.c_void_p]
            L.orc_tree_copy_points.argtypes = [C.c_void_p, _f32p, _f32p, _f32p, _u32p]
            L.orc_expected_num_nodes.restype = C.c_uint64
            L.orc_expected_num_nodes.argtypes = [C.c_uint64, C.c_int, C.c_int]
            L.orc_tree_query.restype = C.c_int
            L.orc_tree_query.argtypes = [C.c_void_p, _f32p, C.c_uint64, C.c_int, C.c_int, C.c_int, _f32p, _u32p, C.c_void_p]
            L.orc_box_distance.restype = C.c_float
            L.orc_box_distance.argtypes = [_f32p, _f32p, C.c_float]
            L.orc_point_distance.restype = C.c_float
            L.orc_point_distance.argtypes = [_f32p, _f32p, C.c_float]
            cls._lib = L
        return cls._lib

    @classmethod
    def philox_points(cls, n: int, seed: int, boxsize: float = 1.0) -> np.ndarray:
        out = np.empty((n, 3), np.float32)
        cls.lib().orc_philox_points(n, seed, boxsize, out.reshape(-1))
        return out

    @classmethod
    def expected_num_nodes(cls, n: int, leaf_size: int, block: int = 8) -> int:
        return int(cls.lib().orc_expected_num_nodes(n, leaf_size, block))

    class Tree(_TreeBase):
        prefix = "orc_"

        def __init__(self, points, leafsize: int = 64, boxsize=None):
            type(self).lib = Oracle.lib()
            pts = _aos(points)
            status = C.c_int(0)
            box = -1.0 if boxsize is None else float(boxsize)
            h = self.lib.orc_tree_build(pts.reshape(-1), pts.shape[0], leafsize, 8, box, C.byref(status))
            if not h:
                if status.value == 1:
                    raise RuntimeError(
                        "When using periodic boundary conditions, all points must be within the box (0 <= x <= box_size)."
                    )
                raise RuntimeError(f"oracle build failed (status {status.value})")
            super().__init__(h, boxsize is not None, 0.0 if boxsize is None else float(boxsize))

        def query(self, q, k: int = 1, workers: int = 1, brute: bool = False, return_stats: bool = False,
                  squared: bool = False):
            q = _aos(q)
            if k <= 0:
                raise RuntimeError("k must be positive integer")
            m = q.shape[0]
            d = np.empty((m, k), np.float32)
            i = np.empty((m, k), np.uint32)
            stats = np.zeros(3, np.uint64)
            if workers <= 0:
                workers = os.cpu_count() or 1
            self.lib.orc_tree_query(self._h, q.reshape(-1), m, k, workers, int(brute) | (2 if squared else 0),
                                    d.reshape(-1), i.reshape(-1),
                                    stats.ctypes.data_as(C.c_void_p))
            return (d, i, stats) if return_stats else (d, i)


class Reference:
    """The reference's own code (compiled by oracle/Makefile into oracle/_ref/)."""

    _lib = None
    path = os.path.join(_HERE, "_ref", "libnbref.so")

    @classmethod
    def available(cls) -> bool:
        if os.path.exists(cls.path):
            return True
        if os.path.isdir("/root/reference/kdtree/src/cpp"):
            try:
                build_checkers()
            except Exception:
                return False
        return os.path.exists(cls.path)

    @classmethod
    def lib(cls):
        if cls._lib is None:
            if not cls.available():
                raise RuntimeError("oracle/_ref/libnbref.so not built and /root/reference absent")
            L = C.CDLL(cls.path)
            L.ref_last_error.restype = C.c_char_p
            L.ref_hardware_concurrency.restype = C.c_int
            L.ref_philox_points.argtypes = [C.c_uint32, C.c_uint, C.c_float, _f32p]
            L.ref_tree_build.restype = C.c_void_p
            L.ref_tree_build.argtypes = [_f32p, C.c_uint64, C.c_int, C.c_float, C.POINTER(C.c_double)]
            L.ref_tree_free.argtypes = [C.c_void_p]
            L.ref_tree_num_points.restype = C.c_uint64
            L.ref_tree_num_points.argtypes = [C.c_void_p]
            L.ref_tree_num_nodes.restype = C.c_uint64
            L.ref_tree_num_nodes.argtypes = [C.c_void_p]
            L.ref_tree_copy_nodes.argtypes = [C.c_void_p, C.c_void_p]
            L.ref_tree_copy_points.argtypes = [C.c_void_p, _f32p, _f32p, _f32p, _u32p]
            L.ref_tree_query.restype = C.c_int
            L.ref_tree_query.argtypes = [C.c_void_p, _f32p, C.c_uint64, C.c_int, C.c_int, _f32p, _u32p, C.c_void_p]
            L.ref_tree_query_ex.restype = C.c_int
            L.ref_tree_query_ex.argtypes = [C.c_void_p, _f32p, C.c_uint64, C.c_int, C.c_int, C.c_int, _f32p, _u32p,
                                            C.c_void_p]
            L.ref_box_distance.restype = C.c_float
            L.ref_box_distance.argtypes = [_f32p, _f32p, C.c_float]
            L.ref_point_distance.restype = C.c_float
            L.ref_point_distance.argtypes = [_f32p, _f32p, C.c_float]
            cls._lib = L
        return cls._lib

    @classmethod
    def philox_points(cls, n: int, seed: int, boxsize: float = 1.0) -> np.ndarray:
        out = np.empty((n, 3), np.float32)
        cls.lib().ref_philox_points(n, seed, boxsize, out.reshape(-1))
        return out

    @classmethod
    def hardware_concurrency(cls) -> int:
        return int(cls.lib().ref_hardware_concurrency())

    class Tree(_TreeBase):
        prefix = "ref_"

        def __init__(self, points, leafsize: int = 64, boxsize=None):
            type(self).lib = Reference.lib()
            pts = _aos(points)
            secs = C.c_double(0.0)
            box = -1.0 if boxsize is None else float(boxsize)
            h = self.lib.ref_tree_build(pts.reshape(-1), pts.shape[0], leafsize, box, C.byref(secs))
            if not h:
                raise RuntimeError(self.lib.ref_last_error().decode())
            self.build_seconds = secs.value
            super().__init__(h, boxsize is not None, 0.0 if boxsize is None else float(boxsize))

        def query(self, q, k: int = 1, workers: int = 1, return_stats: bool = False, squared: bool = False):
            q = _aos(q)
            m = q.shape[0]
            d = np.empty((m, max(k, 0)), np.float32)
            i = np.empty((m, max(k, 0)), np.uint32)
            stats = np.zeros(3, np.uint64)
            rc = self.lib.ref_tree_query_ex(self._h, q.reshape(-1), m, k, workers, int(squared), d.reshape(-1),
                                            i.reshape(-1), stats.ctypes.data_as(C.c_void_p))
            if rc:
                raise RuntimeError(self.lib.ref_last_error().decode())
            return (d, i, stats) if return_stats else (d, i)


# ------------------------------------------------------------------------------------------------
# Parity checker
# ------------------------------------------------------------------------------------------------
def point_distance(points: np.ndarray, q: np.ndarray, boxsize=None, squared: bool = False) -> np.ndarray:
    """Reference-arithmetic Euclidean distance (float32, FMA-free; kdtree.hpp:22-31,71-84) of
    every row of ``points`` to the single query ``q`` -- numpy float32 ops round like the C code.
    ``squared``: the value before postprocess()."""
    p = np.asarray(points, np.float32)
    q = np.asarray(q, np.float32)
    acc = np.zeros(p.shape[0], np.float32)
    for a in range(3):
        d = p[:, a] - q[a]
        if boxsize is None:
            t = d * d
        else:
            L = np.float32(boxsize)
            dp, dm = d + L, d - L
            t = np.minimum(np.minimum(d * d, dp * dp), dm * dm)
        acc = acc + t
    return acc if squared else np.sqrt(acc)


@dataclass
class ParityReport:
    rows: int
    rows_equal: int  # distances bit-equal and indices equal as returned
    rows_equal_after_tie_canonicalisation: int  # equal once rows are ordered by (distance, index)
    rows_boundary_tie_verified: int  # index sets differ only by points exactly tied at the k-th distance
    rows_wrong: int
    first_wrong: int = -1

    @property
    def ok(self) -> bool:
        return self.rows_wrong == 0


def _canon(d: np.ndarray, i: np.ndarray):
    order = np.lexsort((i, d), axis=-1) if d.ndim == 1 else np.stack(
        [np.lexsort((i[r], d[r])) for r in range(d.shape[0])])
    if d.ndim == 1:
        return d[order], i[order]
    return np.take_along_axis(d, order, 1), np.take_along_axis(i, order, 1)


def compare_knn(d_test, i_test, d_ref, i_ref, points=None, queries=None, boxsize=None,
                squared: bool = False) -> ParityReport:
    """Compares a (distances, indices) result with the reference's on the same inputs.

    Distances must be bit-equal row by row.  Index rows must be equal, or equal after ordering
    exact ties by index (the reference leaves the order of equal distances unspecified:
    kdtree_opt.hpp:13-18), or differ only in points whose recomputed distance is exactly the k-th
    distance (the reference keeps the first-visited of such points: kdtree_asm_systemv.asm:155-169);
    the latter needs ``points``/``queries`` for the recomputation.  ``squared``: both results hold
    squared distances (the north star's bar: d2 bit-exact), the recomputation then skips the sqrt too.
    """
    d_test = np.asarray(d_test, np.float32)
    d_ref = np.asarray(d_ref, np.float32)
    i_test = np.asarray(i_test, np.uint32)
    i_ref = np.asarray(i_ref, np.uint32)
    assert d_test.shape == d_ref.shape == i_test.shape == i_ref.shape
    m = d_test.shape[0]
    d_bits_equal = (d_test.view(np.uint32) == d_ref.view(np.uint32)).all(axis=1)
    same = d_bits_equal & (i_test == i_ref).all(axis=1)
    rep = ParityReport(m, int(same.sum()), 0, 0, 0)
    for r in np.nonzero(~same)[0]:
        dt, it = _canon(d_test[r], i_test[r])
        dr, ir = _canon(d_ref[r], i_ref[r])
        if not np.array_equal(dt.view(np.uint32), dr.view(np.uint32)):
            rep.rows_wrong += 1
        elif np.array_equal(it, ir):
            rep.rows_equal_after_tie_canonicalisation += 1
        else:
            ok = False
            if points is not None and queries is not None:
                kth = dt[-1]
                diff = np.setxor1d(it, ir)
                diff = diff[diff < np.asarray(points).shape[0]]
                dd = point_distance(np.asarray(points)[diff], np.asarray(queries)[r], boxsize, squared)
                inner_t = np.sort(it[dt < kth])
                inner_r = np.sort(ir[dr < kth])
                ok = diff.size > 0 and bool((dd.view(np.uint32) == kth.view(np.uint32)).all()) and np.array_equal(inner_t, inner_r)
            if ok:
                rep.rows_boundary_tie_verified += 1
            else:
                rep.rows_wrong += 1
        if rep.rows_wrong == 1 and rep.first_wrong < 0:
            rep.first_wrong = int(r)
    return rep
