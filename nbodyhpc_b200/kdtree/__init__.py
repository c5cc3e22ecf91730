"""Drop-in for ``nbodyhpc.kdtree`` (reference: kdtree/src/python/nbodyhpc/kdtree/__init__.py:11-56).

Same class name, constructor and ``query`` signatures, defaults (``leafsize=128`` here, 64 in the
C++/pybind layer), return types (float32 distances, uint32 indices, both ``(M, k)``) and warning on
unknown keyword arguments.  The tree is built and searched on a B200; there is no CPU fallback and
importing this module fails loudly if the native extension has not been built.

Extensions beyond the reference, all opt-in:

* points / queries may be DEVICE arrays -- anything exposing ``__cuda_array_interface__`` (torch CUDA
  tensors, cupy, numba): they are used in place, no host round trip; query results then come back on
  the device (torch tensors for torch inputs, otherwise ``DeviceArray`` objects that again expose
  ``__cuda_array_interface__``);
* ``query_cdf`` -- histograms of the k-th neighbour distance without materialising the ``(M, k)``
  rows (the kNN-CDF use case);
* N-d query arrays are reshaped as the reference evidently intended (its own reshape raises);
* ``query(..., return_squared=True)`` returns the squared distances the search ranks by (no ``sqrt``);
* any ``k`` is accepted, like the reference's queue (k > 64 keeps the candidates in device memory).
"""
from __future__ import annotations

import warnings
from typing import Optional, Sequence, Tuple

import numpy as np

try:
    from ._impl import KDTree as cKDTree
except ImportError as exc:  # pragma: no cover - build problem, never a silent fallback
    raise ImportError(
        "nbodyhpc_b200.kdtree._impl is not built; run `python -m nbodyhpc_b200._build` "
        "(needs nvcc + pybind11). There is no pure-Python or CPU fallback."
    ) from exc


def _is_device_array(a) -> bool:
    return hasattr(a, "__cuda_array_interface__") and not isinstance(a, np.ndarray)


def _device_view(a, what: str):
    """(pointer, shape) of a C-contiguous float32 device array; raises otherwise (no silent copies)."""
    cai = a.__cuda_array_interface__
    if np.dtype(cai["typestr"]) != np.float32:
        raise TypeError(f"{what}: device arrays must be float32 (got {cai['typestr']}); cast on the device first")
    shape = tuple(int(s) for s in cai["shape"])
    strides = cai.get("strides")
    if strides is not None:
        expect, acc = [], 4
        for s in reversed(shape):
            expect.append(acc)
            acc *= max(s, 1)
        if tuple(int(s) for s in strides) != tuple(reversed(expect)) and int(np.prod(shape)) > 0:
            raise TypeError(f"{what}: device arrays must be C-contiguous")
    return int(cai["data"][0]), shape


def _is_torch(a) -> bool:
    return type(a).__module__.split(".")[0] == "torch"


class DeviceArray:
    """A result left on the device for non-torch callers: owns the memory, exposes
    ``__cuda_array_interface__`` (cupy.asarray / numba.cuda.as_cuda_array / torch.as_tensor accept it)."""

    def __init__(self, shape, dtype, device: int = -1):
        from .. import capi

        self.shape, self.dtype = tuple(shape), np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        self._ptr = capi.device_alloc(self.nbytes, device)  # on the tree's GPU, not the thread's current one

    @property
    def __cuda_array_interface__(self):
        return {"shape": self.shape, "typestr": self.dtype.str, "data": (self._ptr, False), "version": 3,
                "strides": None}

    def to_host(self) -> np.ndarray:
        from .. import capi

        out = np.empty(self.shape, self.dtype)
        capi.device_to_host(out, self._ptr)
        return out

    def __del__(self):
        try:
            from .. import capi

            if self._ptr:
                capi.device_free(self._ptr)
                self._ptr = 0
        except Exception:
            pass


class KDTree(cKDTree):
    """Spatial KD-tree (3-D) with optional periodic boundary conditions, resident on one GPU."""

    def __init__(self, points, leafsize: int = 128, max_threads: int = -1,
                 boxsize: Optional[float] = None, **kwargs):
        """Build a new KDTree.

        Parameters
        ----------
        points : (N, 3) array; copied (and cast to float32) into device memory.  A device array
            (``__cuda_array_interface__``, float32, C-contiguous) is read in place.
        leafsize : points per leaf where the search switches to brute force (effective minimum 16).
        max_threads : accepted for compatibility; construction runs on the GPU.
        boxsize : if not None, the periodic box size; all points must satisfy 0 <= x <= boxsize.
        device : (keyword, extension) CUDA device ordinal; default the current device.
        devices : (keyword, extension) several ordinals: the tree is built on the first and copied
            byte for byte to the others (peer copies); host-array queries are then split into
            contiguous chunks, one per GPU, like the reference's thread pool splits them over threads
            (pybind.cpp:164-172).
        """
        devices = kwargs.pop("devices", None)
        device = kwargs.pop("device", -1)
        if devices is not None:
            devices = [int(d) for d in devices]
            if not devices:
                raise ValueError("devices must name at least one GPU")
            device = devices[0]
        if _is_device_array(points):
            ptr, shape = _device_view(points, "points")
            if len(shape) != 2 or shape[1] != 3:
                raise RuntimeError("positions must be a 2D array of shape (N, 3)")
            if shape[0]:
                from .. import capi

                where = capi.pointer_device(ptr)
                if device == -1:
                    device = where  # the tree is built where its points are
                elif where != device:
                    raise RuntimeError(f"points live on GPU {where} but device={device} was requested")
            stream = _current_stream(points)
            super().__init__(ptr, shape[0], leafsize, max_threads, boxsize, device, stream)
        else:
            super().__init__(points, leafsize, max_threads, boxsize, device)
        self._shards = None
        if devices is not None and len(devices) > 1:
            from .. import capi

            mine = capi.Tree(self._handle, owned=False)
            self._shards = [mine] + [mine.clone_to_device(d) for d in devices[1:]]
        if len(kwargs) > 0:
            warnings.warn("Unrecognized keyword arguments: {}".format(kwargs))

    def _query_sharded(self, points: np.ndarray, k: int, squared: bool = False):
        """Contiguous chunks of the batch, one per replica, answered concurrently."""
        from concurrent.futures import ThreadPoolExecutor

        from ..dist import shard_range

        if k <= 0:
            raise RuntimeError("k must be positive integer")
        points = np.ascontiguousarray(points, dtype=np.float32)
        if points.ndim != 2 or points.shape[1] != 3:
            raise RuntimeError("positions must be a 2D array of shape (N, 3)")
        m, world = points.shape[0], len(self._shards)
        dist, idx = np.empty((m, k), np.float32), np.empty((m, k), np.uint32)

        def run(rank):
            b, e = shard_range(m, rank, world)
            if e > b:
                self._shards[rank].query(points[b:e], k, out=(dist[b:e], idx[b:e]), squared=squared)

        with ThreadPoolExecutor(max_workers=world) as pool:
            list(pool.map(run, range(world)))
        return dist, idx

    def query(self, points, k: int = 1, workers: int = 1, **kwargs) -> Tuple[np.ndarray, np.ndarray]:
        """k nearest neighbours of every query point: ``(distances, indices)`` of shape ``(..., k)``.

        ``workers`` is accepted for compatibility (the batch is one GPU launch sequence).  Device
        arrays in give device arrays out, enqueued on the caller's current stream.
        ``return_squared=True`` (extension) skips the final ``sqrt``: the distances are then the float32
        squared distances of the reference's metric, bit for bit.
        """
        squared = bool(kwargs.pop("return_squared", False))
        if len(kwargs) > 0:
            warnings.warn("Unrecognized keyword arguments: {}".format(kwargs))

        if _is_device_array(points):
            return self._query_device_array(points, k, squared)

        points = np.asarray(points)
        if points.ndim != 2:
            shape = points.shape
            points = points.reshape((-1, shape[-1]))
        else:
            shape = None

        if self._shards is not None:
            distances, indices = self._query_sharded(points, k, squared)
        else:
            distances, indices = super().query(points, k, workers, squared)

        if shape is not None:
            # the reference passes (shape[:-1], k) to reshape, which raises TypeError
            # (__init__.py:52-54); this is the evidently intended result
            distances = distances.reshape(shape[:-1] + (k,))
            indices = indices.reshape(shape[:-1] + (k,))

        return distances, indices

    def _check_same_device(self, ptr: int, m: int):
        """Device arrays are used in place, so they must live on the tree's GPU."""
        if m == 0:
            return
        from .. import capi

        where = capi.pointer_device(ptr)
        if where != self.device:
            raise RuntimeError(f"device array lives on GPU {where}, the tree on GPU {self.device}; "
                               "move the array (or build the tree with device=...)")

    def _query_device_array(self, points, k: int, squared: bool = False):
        if k <= 0:
            raise RuntimeError("k must be positive integer")
        ptr, shape = _device_view(points, "points")
        if len(shape) < 1 or shape[-1] != 3:
            raise RuntimeError("positions must be a 2D array of shape (N, 3)")
        m = int(np.prod(shape[:-1]))
        self._check_same_device(ptr, m)
        out_shape = shape[:-1] + (k,)
        stream = _current_stream(points)
        if _is_torch(points):
            import torch

            dist = torch.empty(out_shape, dtype=torch.float32, device=points.device)
            idx = torch.empty(out_shape, dtype=torch.int32, device=points.device)  # uint32 bit patterns
            if m:
                self._query_device(ptr, m, k, dist.data_ptr(), idx.data_ptr(), stream, squared)
            return dist, idx
        dist, idx = DeviceArray(out_shape, np.float32, self.device), DeviceArray(out_shape, np.uint32, self.device)
        if m:
            self._query_device(ptr, m, k, dist._ptr, idx._ptr, stream, squared)
        return dist, idx

    def query_cdf(self, points, ks: Sequence[int], bins) -> np.ndarray:
        """Histograms of the distance to the k-th neighbour for every k in ``ks`` (extension).

        Returns ``counts`` of shape ``(len(ks), len(bins) - 1)`` with
        ``counts[i] == numpy.histogram(self.query(points, max(ks))[0][:, ks[i] - 1], bins)[0]``
        (``bins`` = increasing float32 edges, last bin closed), computed in ONE traversal at
        ``k = max(ks)`` with the histogram accumulated on the device: nothing per query is written.
        """
        ks = [int(k) for k in ks]
        edges = np.ascontiguousarray(bins, dtype=np.float32)
        if edges.ndim != 1 or len(edges) < 2:
            raise RuntimeError("edges must be a 1D array of at least 2 values")
        if _is_device_array(points):
            from .. import capi

            ptr, shape = _device_view(points, "points")
            if len(shape) < 1 or shape[-1] != 3:
                raise RuntimeError("positions must be a 2D array of shape (N, 3)")
            m = int(np.prod(shape[:-1]))
            self._check_same_device(ptr, m)
            # both helpers complete before they return, so the kernels on the caller's stream see them
            d_edges = DeviceArray(edges.shape, np.float32, self.device)
            capi.host_to_device(d_edges._ptr, edges)
            d_counts = DeviceArray((len(ks), len(edges) - 1), np.uint64, self.device)
            capi.device_memset(d_counts._ptr, d_counts.nbytes)
            self._knn_cdf_device(ptr, m, ks, d_edges._ptr, len(edges) - 1, d_counts._ptr, _current_stream(points))
            return d_counts.to_host()
        points = np.asarray(points, dtype=np.float32).reshape((-1, 3))
        return self._knn_cdf(points, ks, edges)

    def nodes(self) -> np.ndarray:
        """Node records {dim, split, left, right} in the reference's pre-order (extension)."""
        from ..capi import NODE_DTYPE

        return self._nodes_bytes().view(NODE_DTYPE)


def _current_stream(a) -> int:
    """The stream the caller's framework is working on, so that our kernels are ordered after the
    producer of `a` (torch: current stream; otherwise the legacy default stream)."""
    if _is_torch(a):
        import torch

        return int(torch.cuda.current_stream(a.device).cuda_stream)
    return 0


__all__ = ["KDTree", "DeviceArray"]
