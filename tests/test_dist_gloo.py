"""world_size-2 `gloo` test of the multi-GPU host logic: the tree description travels with one
broadcast, queries are sharded in contiguous chunks, and the gathered per-rank answers equal the
single-process answer.  The per-rank "device" here is the CPU oracle (test infrastructure); on the
GPU box the same helpers drive one B200 per rank (tests/test_gpu_multi.py, bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import Oracle, philox


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nbodyhpc_b200 import capi
    from nbodyhpc_b200.dist import broadcast_meta, shard_range

    meta = None
    if rank == 0:
        meta = capi.TreeMeta()
        meta.n_points, meta.n_padded, meta.n_nodes, meta.arena_bytes = 1001, 1008, 31, 123456
        meta.leaf_size, meta.block_size, meta.periodic, meta.box_size = 64, 8, 1, 2.0
        meta.lo[:] = [0.0, 0.1, 0.2]
        meta.hi[:] = [1.0, 1.1, 1.2]
        meta.n_levels = 4
    got = broadcast_meta(meta, src=0)
    assert (got.n_points, got.n_padded, got.n_nodes, got.arena_bytes) == (1001, 1008, 31, 123456)
    assert (got.leaf_size, got.periodic, got.box_size, got.n_levels) == (64, 1, 2.0, 4)
    assert list(got.hi) == pytest.approx([1.0, 1.1, 1.2])

    # replicate-by-value (every rank rebuilds the same deterministic oracle tree) + shard the queries
    pts, q = philox(5000, 42, 2.0), philox(1001, 43, 2.0)
    tree = Oracle.Tree(pts, 64, 2.0)
    b, e = shard_range(len(q), rank, world)
    d, i = tree.query(q[b:e], 4)
    # the sharded kNN-CDF: per-rank histograms of the 4th-neighbour distance, summed by one all-reduce
    from nbodyhpc_b200.dist import reduce_cdf_counts

    edges = np.linspace(0.0, 0.5, 11).astype(np.float32)
    local = np.stack([np.histogram(d[:, k - 1], bins=edges)[0] for k in (1, 4)]).astype(np.uint64)
    total = reduce_cdf_counts(local)
    np.savez(os.path.join(out_dir, f"shard{rank}.npz"), d=d, i=i, b=b, e=e, cdf=total)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_query_equals_single_process(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    pts, q = philox(5000, 42, 2.0), philox(1001, 43, 2.0)
    d_full, i_full = Oracle.Tree(pts, 64, 2.0).query(q, 4)
    shards = [np.load(tmp_path / f"shard{r}.npz") for r in range(world)]
    assert shards[0]["b"] == 0 and shards[0]["e"] == shards[1]["b"] and shards[1]["e"] == len(q)
    d = np.concatenate([s["d"] for s in shards])
    i = np.concatenate([s["i"] for s in shards])
    assert np.array_equal(d.view(np.uint32), d_full.view(np.uint32)) and np.array_equal(i, i_full)
    edges = np.linspace(0.0, 0.5, 11).astype(np.float32)
    expect = np.stack([np.histogram(d_full[:, k - 1], bins=edges)[0] for k in (1, 4)]).astype(np.uint64)
    for s in shards:  # every rank holds the total
        assert np.array_equal(s["cdf"], expect)
