"""The Q = 100 M point of the C5 sweep (10 M-point surface cloud, k = 16, ~80 queries per cell) with each block-kernel variant of this
process's environment: default = knn_thr_kernel (per-lane walk); PCC_THR_STAGED=1 = TMA-staged tiles; PCC_CELL_KERNEL=1 = round 1's
warp-owns-a-cell TMA kernel; PCC_OLD_FAST=1 = round 1's block kernel.  Parity of 20 000 sampled rows against the oracle each time."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle
from pointcloudcomparator_b200 import synth
from pointcloudcomparator_b200.search import GridSearch
N, Q = 10_000_000, int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
cache = f"/tmp/pcc_probe_surface_{N}.npz"
if os.path.exists(cache): ref = np.load(cache)["ref"]
else: ref = synth.room(N, 4001, size=(10, 10, 3), stride4=True)
dref = torch.from_numpy(ref).cuda()
g = torch.Generator(device="cuda"); g.manual_seed(5002)
pick = torch.randint(0, N, (Q,), device="cuda", generator=g)
dq = dref[pick].clone(); dq[:, :3] += torch.randn((Q, 3), device="cuda", generator=g) * 0.01
s = GridSearch(0).setInputCloud(dref, k_hint=16); s.setTiming(True)
best = 1e9
for _ in range(3):
    idx, d2, _ = s.nearestKSearch(dq, 16); best = min(best, s.lastKernelMs())
sel = torch.randint(0, Q, (20000,), device="cuda", generator=g)
oi, od, _ = oracle.KdTree(ref).knn(dq[sel].cpu().numpy(), 16)
ok = bool(np.array_equal(idx[sel].cpu().numpy(), oi) and np.array_equal(d2[sel].cpu().numpy().view(np.uint32), od.view(np.uint32)))
env = {k: os.environ[k] for k in ("PCC_THR_STAGED", "PCC_CELL_KERNEL", "PCC_OLD_FAST") if k in os.environ}
print(json.dumps(dict(variant=env or "default (knn_thr_kernel)", n_ref=N, n_query=Q, k=16, kernel_ms=round(best, 3), gqps=round(Q / best / 1e6, 3), parity_20000_rows=ok)), flush=True)
