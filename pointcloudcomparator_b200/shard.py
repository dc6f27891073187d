"""Query sharding across the GPUs of one node.

The path shards by construction: queries are independent given an immutable reference grid.  One process
per GPU (torch.distributed); the reference grid is built ONCE on rank 0 and broadcast (NCCL over NVLink) as
its two flat device arrays; every rank then answers its own contiguous, cell-ordered range of the queries
with no data-path collective.  A gather of the per-shard results and the min-label merge of Euclidean
cluster labels are the only other collectives, and both are optional (consumers can stay sharded).

Host logic here is backend-agnostic so it is covered by world_size-2 gloo tests on CPU (tests/test_shard.py);
only `broadcast_grid` touches the CUDA library.
"""
from __future__ import annotations

import numpy as np

try:
    import torch
    import torch.distributed as dist
except Exception:  # pragma: no cover
    torch = None
    dist = None


def cell_keys(points: np.ndarray, origin, cell: float, dims) -> np.ndarray:
    """Row-major (z, y, x) cell key of each point -- the order libpcc_search sorts the grid in."""
    p = np.asarray(points, np.float32)[:, :3]
    inv = np.float32(1.0) / np.float32(cell)
    u = (p - np.asarray(origin, np.float32)) * inv
    c = np.clip(np.floor(u), 0, np.asarray(dims, np.float32) - 1).astype(np.int64)
    return (c[:, 2] * dims[1] + c[:, 1]) * dims[0] + c[:, 0]


def shard_ranges(n: int, world: int):
    """Contiguous equal-count [begin, end) ranges; the first n % world shards get one extra query."""
    base, extra = divmod(int(n), int(world))
    out, b = [], 0
    for r in range(world):
        e = b + base + (1 if r < extra else 0)
        out.append((b, e))
        b = e
    return out


def shard_queries(queries: np.ndarray, rank: int, world: int, origin=None, cell: float | None = None, dims=None, block: int = 0):
    """This rank's share of `queries` and the row numbers it came from.

    With a grid description the queries are first ordered by cell key so each shard is made of compact slabs of the
    reference grid (good L2 reuse); without one the split is by input order.
    block = 0: ONE contiguous range of the cell order per rank.  block > 0: the cell order is cut into blocks of that many
    queries, dealt to the ranks round-robin -- every rank gets the same number of queries (+-1 block) AND the same mix of
    easy (flat floor) and hard (edges, clutter) regions; measured on 8 B200s the contiguous split left the slowest rank 40 %
    behind the fastest although the counts were equal.  Rows stay in cell order inside a shard either way.
    """
    q = np.asarray(queries)
    if origin is not None:
        order = np.argsort(cell_keys(q, origin, cell, dims), kind="stable")
    else:
        order = np.arange(q.shape[0])
    if block and block > 0 and world > 1:
        nblk = (q.shape[0] + block - 1) // block
        mine = [order[b * block:(b + 1) * block] for b in range(rank, nblk, world)]
        rows = np.concatenate(mine) if mine else order[:0]
    else:
        b, e = shard_ranges(q.shape[0], world)[rank]
        rows = order[b:e]
    return np.ascontiguousarray(q[rows]), rows


def gather_rows(local: "torch.Tensor", rows: "torch.Tensor", n_total: int, group=None):
    """All-gather per-shard result rows back into original query order (every rank gets the full table)."""
    world = dist.get_world_size(group)
    counts = [torch.zeros(1, dtype=torch.int64, device=local.device) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device), group=group)
    counts = [int(c.item()) for c in counts]
    mx = max(counts)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    padr = torch.zeros(mx, dtype=torch.int64, device=local.device)
    padr[: rows.shape[0]] = rows
    parts = [torch.empty_like(pad) for _ in range(world)]
    rparts = [torch.empty_like(padr) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    dist.all_gather(rparts, padr, group=group)
    out = torch.empty((n_total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for p, r, c in zip(parts, rparts, counts):
        out[r[:c]] = p[:c]
    return out


def allreduce_sums(sums: np.ndarray, count: int, device=None, group=None):
    """ICP: the 16 correspondence sums + count of every shard added in fp64 (the only exchange an ICP step needs)."""
    t = torch.tensor(list(sums) + [float(count)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t[:16].cpu().numpy(), int(round(float(t[16])))


def merge_labels_min(labels: "torch.Tensor", edges_a: "torch.Tensor", edges_b: "torch.Tensor", group=None, max_rounds: int = 64):
    """Euclidean clustering across shards: every rank holds a full-length component-label array in which it hooked
    only the edges whose query endpoint it owns.  Labels converge by repeating {all-reduce(min), re-apply local
    edges, pointer-jump} until no label changes anywhere.  Labels are point indices (a component's label = its
    smallest member)."""
    lab = labels.clone()
    for _ in range(max_rounds):
        before = lab.clone()
        dist.all_reduce(lab, op=dist.ReduceOp.MIN, group=group)
        for _ in range(64):                         # local fix-point: edges + pointer jumping
            prev = lab.clone()
            if edges_a.numel():
                m = torch.minimum(lab[edges_a], lab[edges_b])
                lab.scatter_reduce_(0, edges_a, m, reduce="amin")
                lab.scatter_reduce_(0, edges_b, m, reduce="amin")
            lab = torch.minimum(lab, lab[lab])
            if torch.equal(prev, lab):
                break
        changed = torch.tensor([0 if torch.equal(before, lab) else 1], dtype=torch.int32, device=lab.device)
        dist.all_reduce(changed, op=dist.ReduceOp.MAX, group=group)
        if int(changed.item()) == 0:
            break
    return lab


def euclidean_clusters_sharded(search, tolerance: float, min_size: int, max_size: int, group=None, max_rounds: int = 64):
    """EuclideanClusterExtraction with the radius-graph work sharded over the ranks (replicated grid).

    Each rank links only the edges whose query endpoint lies in its contiguous range of the sorted order
    (`pcc_ece_link_range`), then the ranks repeat {element-wise MIN all-reduce of the compressed forests (NCCL),
    absorb the result (`pcc_ece_absorb`)} until the all-reduce is a fix-point -- typically 2-3 rounds.  Every rank ends
    with the same forest and finishes locally (sizes, size filter, PCL ordering).  Returns (labels[n_input], sizes)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    b, e = shard_ranges(search.size, world)[rank]
    parent = search.eceNewForest()
    search.eceLinkRange(parent, tolerance, b, e)
    search.eceAbsorb(parent, parent)                       # compress: parent[i] = root = smallest member
    for _ in range(max_rounds):
        merged = parent.clone()
        dist.all_reduce(merged, op=dist.ReduceOp.MIN, group=group)
        changed = torch.tensor([0 if torch.equal(merged, parent) else 1], dtype=torch.int32, device=parent.device)
        dist.all_reduce(changed, op=dist.ReduceOp.MAX, group=group)
        if int(changed.item()) == 0:
            break
        search.eceAbsorb(parent, merged)
    return search.eceFinish(parent, min_size, max_size)


# ---- C-ABI multi-GPU path: an NCCL communicator of our own, handed to libpcc_search (pcc_comm_init) ----------------------------
_NCCL = None


def _nccl():
    """libnccl.so.2 as the process already has it (torch's bundled copy once torch is imported); libpcc_search dlopens the same one."""
    import ctypes
    global _NCCL
    if _NCCL is None:
        _NCCL = ctypes.CDLL("libnccl.so.2", mode=ctypes.RTLD_GLOBAL)
        _NCCL.ncclGetErrorString.restype = ctypes.c_char_p
    return _NCCL


def nccl_comm(device: int, group=None) -> int:
    """ncclCommInitRank over the ranks of `group` (the 128-byte unique id travels through torch.distributed).  Returns the
    ncclComm_t as an integer; pass it to GridSearch.commInit.  The communicator lives until the process exits."""
    import ctypes

    class _Uid(ctypes.Structure):
        _fields_ = [("internal", ctypes.c_char * 128)]

    lib = _nccl()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    uid = _Uid()
    if rank == 0:
        rc = lib.ncclGetUniqueId(ctypes.byref(uid))
        if rc != 0:
            raise RuntimeError(f"ncclGetUniqueId: {lib.ncclGetErrorString(rc).decode()}")
    box = [bytes(uid) if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    uid = _Uid.from_buffer_copy(box[0])
    torch.cuda.set_device(device)
    comm = ctypes.c_void_p()
    lib.ncclCommInitRank.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, _Uid, ctypes.c_int]
    rc = lib.ncclCommInitRank(ctypes.byref(comm), world, uid, rank)
    if rc != 0:
        raise RuntimeError(f"ncclCommInitRank: {lib.ncclGetErrorString(rc).decode()}")
    return int(comm.value)


def attach(search, group=None):
    """Create a communicator over `group` and attach it to `search` (pcc_comm_init).  Returns (rank, world)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    search.commInit(nccl_comm(search.device, group), rank, world)
    return rank, world


def broadcast_grid(search, src: int = 0, group=None):
    """Rank `src` has a built GridSearch; every other rank adopts its grid (meta via broadcast_object_list,
    the float4 points and the cell-start table via NCCL broadcast straight into the adopted device arrays)."""
    rank = dist.get_rank(group)
    dev = torch.device(f"cuda:{search.device}")
    if rank == src:
        meta, p_pts, p_cells = search.export()
        box = [meta.tolist()]
    else:
        box = [None]
    dist.broadcast_object_list(box, src=src, group=group)
    meta = np.asarray(box[0], np.float64)
    if rank != src:
        _, p_pts, p_cells = search.adopt(meta)
    n_pts, n_cells = int(meta[0]), int(meta[11])

    def view(ptr, nbytes):
        # zero-copy torch view (float32 words) of library-owned device memory
        class _Mem:
            pass
        m = _Mem()
        m.__cuda_array_interface__ = {"shape": (max(nbytes // 4, 1),), "typestr": "<f4", "data": (int(ptr), False), "version": 3}
        return torch.as_tensor(m, device=dev)

    if n_pts > 0:
        dist.broadcast(view(p_pts, n_pts * 16), src=src, group=group)
    dist.broadcast(view(p_cells, (n_cells + 1) * 4), src=src, group=group)
    torch.cuda.synchronize(dev)
    return meta
