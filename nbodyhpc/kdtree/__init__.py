"""`nbodyhpc.kdtree` import path of the reference, served by nbodyhpc_b200."""
from nbodyhpc_b200.kdtree import KDTree  # noqa: F401

__all__ = ["KDTree"]
