"""nbodyhpc_b200 -- B200-native (sm_100a) kd-tree build + batched kNN query, a drop-in for the hot
path of wendazhou/nbodyhpc's ``nbodyhpc.kdtree`` package.

    from nbodyhpc_b200.kdtree import KDTree      # same API as nbodyhpc.kdtree.KDTree

Layout: ``csrc/`` CUDA kernels + the C ABI (``include/nbk.h`` -> ``lib/libnbk.so``), ``kdtree/`` the
Python/pybind11 mirror of the reference interface, ``capi`` a ctypes view of the C ABI for device
pointers, ``dist`` the one-process-per-GPU replicate/shard helpers.
"""

__version__ = "0.1.0"
