"""Multi-GPU path on real devices: NCCL-broadcast replica answers byte-identically to the source
tree.  Needs >= 2 GPUs (skipped otherwise); the host logic is covered on CPU by test_dist_gloo.py."""
import os
import socket

import numpy as np
import pytest

from helpers import philox

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from nbodyhpc_b200 import capi
    from nbodyhpc_b200.dist import replicate_tree, shard_range

    pts, q = philox(300_000, 42), philox(40_001, 43)
    tree = capi.Tree.build(pts, 64, 1.0, device=rank) if rank == 0 else None
    tree = replicate_tree(tree, src=0, device=rank)
    b, e = shard_range(len(q), rank, world)
    d, i = tree.query(q[b:e], 8)
    from nbodyhpc_b200.dist import knn_cdf_sharded

    cdf = knn_cdf_sharded(tree, q, [1, 8], np.linspace(0.0, 0.1, 21).astype(np.float32))
    np.savez(os.path.join(out_dir, f"shard{rank}.npz"), d=d, i=i, nodes=tree.nodes().view(np.uint8), cdf=cdf)
    dist.barrier()
    dist.destroy_process_group()


def test_replica_is_byte_identical_and_shards_compose(gpu, tmp_path):
    import torch
    import torch.multiprocessing as mp

    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    pts, q = philox(300_000, 42), philox(40_001, 43)
    tree = gpu.Tree.build(pts, 64, 1.0, device=0)
    d_full, i_full = tree.query(q, 8)
    shards = [np.load(tmp_path / f"shard{r}.npz") for r in range(world)]
    for s in shards:
        assert np.array_equal(s["nodes"], tree.nodes().view(np.uint8))
    d = np.concatenate([s["d"] for s in shards])
    i = np.concatenate([s["i"] for s in shards])
    assert np.array_equal(d.view(np.uint32), d_full.view(np.uint32)) and np.array_equal(i, i_full)
    edges = np.linspace(0.0, 0.1, 21).astype(np.float32)
    expect = np.stack([np.histogram(d_full[:, k - 1], bins=edges)[0] for k in (1, 8)]).astype(np.uint64)
    for s in shards:
        assert np.array_equal(s["cdf"], expect)


def test_in_process_replicas_from_python(gpu):
    """KDTree(..., devices=[...]): the tree is cloned to the other GPUs with peer copies and a
    host-array query is split into contiguous chunks, one per GPU."""
    import torch

    from nbodyhpc_b200.kdtree import KDTree

    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    pts, q = philox(300_000, 42), philox(100_003, 43)
    single = KDTree(pts, leafsize=64, boxsize=1.0, device=0)
    multi = KDTree(pts, leafsize=64, boxsize=1.0, devices=list(range(world)))
    d1, i1 = single.query(q, k=8)
    d2, i2 = multi.query(q, k=8)
    assert np.array_equal(d1.view(np.uint32), d2.view(np.uint32)) and np.array_equal(i1, i2)
    for shard in multi._shards[1:]:
        assert np.array_equal(shard.nodes(), single.nodes())
    assert [s.device for s in multi._shards] == list(range(world))
