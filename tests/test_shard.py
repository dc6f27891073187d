"""Host-side multi-GPU logic (pointcloudcomparator_b200/shard.py) on CPU: world_size-2 gloo process groups.
The oracle stands in for the per-shard search so the test checks the sharding / gather / merge plumbing itself."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from pointcloudcomparator_b200 import shard, synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _run(fn, world=2):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), fn, ret), nprocs=world, join=True)
    return [ret[r] for r in range(world)]


def test_shard_ranges_cover_everything():
    for n, w in ((10, 3), (7, 8), (0, 2), (1000003, 8)):
        r = shard.shard_ranges(n, w)
        assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        sizes = [e - b for b, e in r]
        assert max(sizes) - min(sizes) <= 1


def test_cell_ordered_shards_are_compact_slabs():
    ref = synth.room(20000, 1001)
    origin, cell, dims = ref.min(0), 0.25, (25, 17, 12)
    parts = [shard.shard_queries(ref, r, 4, origin, cell, dims) for r in range(4)]
    rows = np.concatenate([p[1] for p in parts])
    assert np.array_equal(np.sort(rows), np.arange(len(ref)))                 # a partition of the queries
    keys = [shard.cell_keys(p[0], origin, cell, dims) for p in parts]
    assert all(k0.max() <= k1.min() for k0, k1 in zip(keys, keys[1:]))         # contiguous key ranges -> z-slabs


def _knn_sharded(rank, world):
    ref = synth.room(6000, 1001)
    qry = synth.noisy_copy(ref, 1002, 0.002)[:1501]
    mine, rows = shard.shard_queries(qry, rank, world, ref.min(0), 0.3, (21, 14, 10))
    idx, d2, _ = oracle.KdTree(ref).knn(mine, 8)                               # stand-in for the per-GPU search
    full_i = shard.gather_rows(torch.from_numpy(idx), torch.from_numpy(rows), len(qry))
    full_d = shard.gather_rows(torch.from_numpy(d2), torch.from_numpy(rows), len(qry))
    return full_i.numpy(), full_d.numpy()


def test_gather_restores_query_order_gloo():
    out = _run(_knn_sharded)
    ref = synth.room(6000, 1001)
    qry = synth.noisy_copy(ref, 1002, 0.002)[:1501]
    oi, od, _ = oracle.KdTree(ref).knn(qry, 8)
    for gi, gd in out:
        assert np.array_equal(gi, oi) and np.array_equal(gd.view(np.uint32), od.view(np.uint32))


def _icp_sums(rank, world):
    src, tgt, _ = synth.icp_pair(4000, 4001, size=(3, 3, 2))
    mine, _ = shard.shard_queries(src, rank, world)
    cnt, sums, _, _ = oracle.KdTree(tgt).icp_pass(mine[:, :3])
    return shard.allreduce_sums(sums, cnt)


def test_icp_sums_allreduce_gloo():
    out = _run(_icp_sums)
    src, tgt, _ = synth.icp_pair(4000, 4001, size=(3, 3, 2))
    cnt, sums, _, _ = oracle.KdTree(tgt).icp_pass(src[:, :3])
    for s, c in out:
        assert c == cnt and np.allclose(s, sums, rtol=1e-12)
        assert np.allclose(oracle.umeyama_from_sums(s, c), oracle.umeyama_from_sums(sums, cnt), atol=1e-7)


def _cluster_merge(rank, world):
    pts, _ = synth.scene(6000, 3001, extent=4.0, n_objects=8)
    n = len(pts)
    off, nbr, _ = oracle.KdTree(pts).radius(pts, 0.08)
    owner = np.arange(n) % world                                                 # each rank hooks only the edges it owns
    a = np.repeat(np.arange(n), np.diff(off))
    keep = owner[a] == rank
    lab = shard.merge_labels_min(torch.arange(n, dtype=torch.int64), torch.from_numpy(a[keep]), torch.from_numpy(nbr[keep].astype(np.int64)))
    return lab.numpy()


def test_cluster_label_merge_gloo():
    out = _run(_cluster_merge)
    pts, _ = synth.scene(6000, 3001, extent=4.0, n_objects=8)
    olab, sizes = oracle.KdTree(pts).ece(0.08, 1, 10**9)
    assert np.array_equal(out[0], out[1])
    lab = out[0]
    # same partition as the oracle's connected components, and a component's label is its smallest member
    for c in range(len(sizes)):
        members = np.flatnonzero(olab == c)
        assert (lab[members] == members.min()).all()


def test_block_cyclic_shards_partition_and_balance():
    """shard_queries(block > 0): blocks of the cell order dealt round-robin -- every row lands in exactly one shard, counts differ by at
    most one block, rows stay in cell order inside a shard, and every shard samples the whole key range (the point of the scheme:
    the contiguous split left the slowest of 8 ranks 40 % behind the fastest on the bench cloud)."""
    from pointcloudcomparator_b200 import shard, synth
    q = synth.room(50000, 5)
    origin, cell, dims = q.min(0), 0.05, (130, 90, 60)
    keys = shard.cell_keys(q, origin, cell, dims)
    world, block = 4, 1000
    seen = np.zeros(len(q), np.int64)
    for r in range(world):
        mine, rows = shard.shard_queries(q, r, world, origin, cell, dims, block=block)
        assert np.array_equal(mine, q[rows])
        seen[rows] += 1
        assert abs(len(rows) - len(q) / world) <= block
        k = keys[rows]
        assert (np.diff(k) >= 0).all()                                      # cell order inside the shard
        assert k.min() < np.quantile(keys, 0.1) and k.max() > np.quantile(keys, 0.9)     # spans the whole grid
    assert (seen == 1).all()
    # block = 0 keeps the contiguous split
    a, rows = shard.shard_queries(q, 1, world, origin, cell, dims)
    assert len(rows) == len(q) // world and (np.diff(keys[rows]) >= 0).all()
