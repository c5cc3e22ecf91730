"""One process per GPU: replicate the tree, shard the queries.

The reference's only parallelism is `wenda::thread_pool::parallelize_loop` over contiguous query
chunks against a read-only tree (pybind.cpp:164-172, thread_pool.hpp:147-183).  Here a "worker" is
a rank with its own B200: the tree is built once (rank ``src``) and its packed arena is replicated
with ONE NCCL broadcast over NVLink/NVSwitch; every rank then answers its contiguous chunk of the
queries.  No collective follows the query, so there is nothing to fuse a kernel with.

torch is used for what it is here for: `torch.distributed` and wrapping device memory.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import capi


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous chunk of [0, total) owned by ``rank``: floor(total/world) each, the last rank takes
    the remainder -- the split of thread_pool::parallelize_loop (thread_pool.hpp:163-173)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    block = total // world
    if block == 0:
        # fewer items than ranks: one item per rank until they run out (thread_pool.hpp:165-168)
        return (rank, rank + 1) if rank < total else (total, total)
    begin = rank * block
    end = total if rank == world - 1 else begin + block
    return begin, end


class _DeviceBytes:
    """Exposes a raw device allocation through __cuda_array_interface__ (uint8)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {
            "shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2, "strides": None,
        }


def arena_tensor(tree: capi.Tree):
    """uint8 torch view of the tree's packed arena [nodes | 128-byte point tiles] (no copy)."""
    import torch

    ptr, nbytes = tree.arena()
    return torch.as_tensor(_DeviceBytes(ptr, nbytes), device=f"cuda:{tree.device}")


def _meta_to_tensor(meta: Optional[capi.TreeMeta], device):
    import torch

    raw = np.zeros(C.sizeof(capi.TreeMeta), np.uint8)
    if meta is not None:
        raw[:] = np.frombuffer(bytes(meta), np.uint8)
    return torch.from_numpy(raw).to(device)


def _tensor_to_meta(t) -> capi.TreeMeta:
    return capi.TreeMeta.from_buffer_copy(t.cpu().numpy().tobytes())


def broadcast_meta(meta: Optional[capi.TreeMeta], src: int = 0, group=None, device="cpu") -> capi.TreeMeta:
    """Broadcasts the fixed-size tree description (works on any backend; CPU tensors for gloo)."""
    import torch.distributed as dist

    t = _meta_to_tensor(meta, device)
    dist.broadcast(t, src=src, group=group)
    return _tensor_to_meta(t)


def replicate_tree(tree: Optional[capi.Tree], src: int = 0, group=None, device: Optional[int] = None) -> capi.Tree:
    """Returns this rank's replica of rank ``src``'s tree (``tree`` is ignored on other ranks).

    One broadcast of the meta block, one of the arena bytes; replicas are byte-identical, hence
    results are too."""
    import torch
    import torch.distributed as dist

    rank = dist.get_rank(group)
    if device is None:
        device = torch.cuda.current_device()
    meta = broadcast_meta(tree.meta if rank == src else None, src, group, device=f"cuda:{device}")
    if rank != src:
        tree = capi.Tree.alloc_replica(meta, device)
    dist.broadcast(arena_tensor(tree), src=src, group=group)
    return tree


def reduce_cdf_counts(local_counts: np.ndarray, group=None, device: str = "cpu") -> np.ndarray:
    """Sum of the per-rank kNN-CDF histograms (the one step of the sharded path that has an exchange:
    every rank histograms the k-th neighbour distances of ITS queries, the totals are one all-reduce of
    ``len(ks) * n_bins`` integers).  Works on any backend (CPU tensors for gloo, device for nccl)."""
    import torch
    import torch.distributed as dist

    t = torch.from_numpy(np.ascontiguousarray(local_counts).astype(np.int64)).to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy().astype(np.uint64)


def knn_cdf_sharded(tree: capi.Tree, queries: np.ndarray, ks, edges, group=None) -> np.ndarray:
    """Fused kNN-CDF of a query set that is sharded over the ranks: ``queries`` is the WHOLE set (every
    rank holds it, as with the reference's thread pool), each rank answers its contiguous chunk on its
    own replica, and the histograms are summed.  Every rank returns the total."""
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    b, e = shard_range(len(queries), rank, world)
    local = tree.knn_cdf(queries[b:e], ks, edges)
    return reduce_cdf_counts(local, group, device=f"cuda:{tree.device}" if dist.get_backend(group) == "nccl" else "cpu")
