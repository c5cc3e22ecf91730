"""Drop-in for ``nbodyhpc.kdtree`` (reference: kdtree/src/python/nbodyhpc/kdtree/__init__.py:11-56).

Same class name, constructor and ``query`` signatures, defaults (``leafsize=128`` here, 64 in the
C++/pybind layer), return types (float32 distances, uint32 indices, both ``(M, k)``) and warning on
unknown keyword arguments.  The tree is built and searched on a B200; there is no CPU fallback and
importing this module fails loudly if the native extension has not been built.
"""
from __future__ import annotations

import warnings
from typing import Optional, Tuple

import numpy as np

try:
    from ._impl import KDTree as cKDTree
except ImportError as exc:  # pragma: no cover - build problem, never a silent fallback
    raise ImportError(
        "nbodyhpc_b200.kdtree._impl is not built; run `python -m nbodyhpc_b200._build` "
        "(needs nvcc + pybind11). There is no pure-Python or CPU fallback."
    ) from exc


class KDTree(cKDTree):
    """Spatial KD-tree (3-D) with optional periodic boundary conditions, resident on one GPU."""

    def __init__(self, points: np.ndarray, leafsize: int = 128, max_threads: int = -1,
                 boxsize: Optional[float] = None, **kwargs):
        """Build a new KDTree.

        Parameters
        ----------
        points : (N, 3) array; copied (and cast to float32) into device memory.
        leafsize : points per leaf where the search switches to brute force (effective minimum 16).
        max_threads : accepted for compatibility; construction runs on the GPU.
        boxsize : if not None, the periodic box size; all points must satisfy 0 <= x <= boxsize.
        device : (keyword, extension) CUDA device ordinal; default the current device.
        """
        device = kwargs.pop("device", -1)
        super().__init__(points, leafsize, max_threads, boxsize, device)
        if len(kwargs) > 0:
            warnings.warn("Unrecognized keyword arguments: {}".format(kwargs))

    def query(self, points: np.ndarray, k: int = 1, workers: int = 1, **kwargs) -> Tuple[np.ndarray, np.ndarray]:
        """k nearest neighbours of every query point: ``(distances, indices)`` of shape ``(..., k)``.

        ``workers`` is accepted for compatibility (the batch is one GPU launch sequence).
        """
        if len(kwargs) > 0:
            warnings.warn("Unrecognized keyword arguments: {}".format(kwargs))

        points = np.asarray(points)
        if points.ndim != 2:
            shape = points.shape
            points = points.reshape((-1, shape[-1]))
        else:
            shape = None

        distances, indices = super().query(points, k, workers)

        if shape is not None:
            # the reference passes (shape[:-1], k) to reshape, which raises TypeError
            # (__init__.py:52-54); this is the evidently intended result
            distances = distances.reshape(shape[:-1] + (k,))
            indices = indices.reshape(shape[:-1] + (k,))

        return distances, indices

    def nodes(self) -> np.ndarray:
        """Node records {dim, split, left, right} in the reference's pre-order (extension)."""
        from ..capi import NODE_DTYPE

        return self._nodes_bytes().view(NODE_DTYPE)


__all__ = ["KDTree"]
