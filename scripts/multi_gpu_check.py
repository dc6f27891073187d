"""torchrun --nproc-per-node N scripts/multi_gpu_check.py : exercises the NCCL paths of pointcloudcomparator_b200.shard on real GPUs
(grid broadcast + adopt, query-sharded kNN + gather, sharded Euclidean clustering, ICP sums all-reduce) against the CPU oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import oracle
from pointcloudcomparator_b200 import shard, synth
from pointcloudcomparator_b200.search import GridSearch

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
dev = torch.device(f"cuda:{local}")

# 1. grid built on rank 0 only, broadcast, adopted everywhere
pts, ids = synth.scene(400000, 3001, extent=12.0, n_objects=80)
s = GridSearch(local)
if rank == 0:
    s.setInputCloud(torch.from_numpy(pts).cuda(), cell_hint=0.05)
meta = shard.broadcast_grid(s, src=0)
assert s.size == len(pts)

# 2. query-sharded kNN, gathered back to query order
qry = synth.sweep_queries(pts, 50001, seed=7, sigma=0.01)
g = s.grid_info()
origin = meta[5:8]
mine, rows = shard.shard_queries(qry, rank, world, origin, g["cell"], g["dims"])
idx, d2, _ = s.nearestKSearch(torch.from_numpy(mine).cuda(), 16)
full_i = shard.gather_rows(idx, torch.from_numpy(rows).to(dev), len(qry))
full_d = shard.gather_rows(d2, torch.from_numpy(rows).to(dev), len(qry))
oi, od, _ = oracle.KdTree(pts).knn(qry, 16)
assert np.array_equal(full_i.cpu().numpy(), oi) and np.array_equal(full_d.cpu().numpy().view(np.uint32), od.view(np.uint32)), "sharded kNN mismatch"

# 3. sharded Euclidean clustering (link own range, MIN all-reduce, absorb to a fix-point)
lab, sizes = shard.euclidean_clusters_sharded(s, 0.05, 100, 250000)
olab, osizes = oracle.KdTree(pts).ece(0.05, 100, 250000)
assert np.array_equal(lab.cpu().numpy(), olab) and np.array_equal(sizes.cpu().numpy(), osizes), "sharded clustering mismatch"

# 4. ICP correspondence sums: each rank matches its share of the source, sums all-reduced
src, tgt, T = synth.icp_pair(200000, 4001, size=(5, 5, 3))
t = GridSearch(local)
if rank == 0:
    t.setInputCloud(torch.from_numpy(tgt).cuda(), k_hint=8)
shard.broadcast_grid(t, src=0)
mine, _ = shard.shard_queries(src, rank, world)
cnt, sums, _, _ = t.icpStep(torch.from_numpy(np.ascontiguousarray(mine)).cuda(), None)
sums, cnt = shard.allreduce_sums(sums, cnt, device=dev)
ocnt, osums, _, _ = oracle.KdTree(tgt).icp_pass(src[:, :3])
assert cnt == ocnt and np.allclose(sums, osums, rtol=1e-11, atol=1e-8), "sharded ICP sums mismatch"

# 5. the same through the C ABI's own multi-GPU entry points (pcc_comm_init / pcc_broadcast_index / pcc_gather / sharded pcc_icp_align)
c = GridSearch(local)
shard.attach(c)
if rank == 0:
    c.setInputCloud(torch.from_numpy(pts).cuda(), cell_hint=0.05)
c.broadcastIndex(0)
assert c.size == len(pts) and c.grid_info() == s.grid_info()
mine, rows = shard.shard_queries(qry, rank, world, origin, g["cell"], g["dims"])
ci, cd, _ = c.nearestKSearch(torch.from_numpy(mine).cuda(), 16)
drows = torch.from_numpy(rows).to(dev)
assert np.array_equal(c.gather(ci, drows, len(qry)).cpu().numpy(), oi), "pcc_gather (indices) mismatch"
assert np.array_equal(c.gather(cd, drows, len(qry)).cpu().numpy().view(np.uint32), od.view(np.uint32)), "pcc_gather (d2) mismatch"
cat = c.gather(ci, None, len(qry)).cpu().numpy()                 # rank-order concatenation
b0, e0 = shard.shard_ranges(len(qry), world)[rank]
assert np.array_equal(cat[b0:e0], ci.cpu().numpy())
# sharded ICP: every rank aligns its slice of the source; result must equal the single-GPU alignment up to fp64 summation order
t2 = GridSearch(local)
shard.attach(t2)
if rank == 0:
    t2.setInputCloud(torch.from_numpy(tgt).cuda(), k_hint=8)
t2.broadcastIndex(0)
b1, e1 = shard.shard_ranges(len(src), world)[rank]
r_sh = t2.icpAlign(torch.from_numpy(np.ascontiguousarray(src[b1:e1])).cuda(), 20)
ref_icp = oracle.icp(src, tgt, 20)
assert r_sh["converged"] == ref_icp["converged"] and abs(r_sh["iterations"] - ref_icp["iterations"]) <= 1
assert np.allclose(r_sh["T"], ref_icp["T"], atol=2e-5) and np.isclose(r_sh["fitness"], ref_icp["fitness"], rtol=1e-3)
Tall = [None] * world
dist.all_gather_object(Tall, r_sh["T"].tobytes())
assert all(x == Tall[0] for x in Tall), "ranks disagree on the ICP transform"
dist.barrier()
if rank == 0:
    print(f"MULTI_GPU_CHECK PASS world={world}: broadcast grid, sharded kNN gather, sharded clustering ({len(sizes)} clusters), ICP sums; "
          f"C ABI: pcc_comm_init + pcc_broadcast_index + pcc_gather + sharded pcc_icp_align ({r_sh['iterations']} iterations)", flush=True)
dist.destroy_process_group()
