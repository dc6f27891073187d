"""Developer probe: kNN time on a cloud with strong density variation (points per area ~ 1/r around a scanner)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pointcloudcomparator_b200.search import GridSearch
n = int(sys.argv[1]); k = int(sys.argv[2])
rng = np.random.default_rng(11)
r = rng.uniform(0.5, 20.0, n); th = rng.uniform(0, 2 * np.pi, n)
ref = np.zeros((n, 4), np.float32)
ref[:, 0] = r * np.cos(th); ref[:, 1] = r * np.sin(th); ref[:, 2] = rng.normal(0, 0.002, n) + 0.05 * np.sin(ref[:, 0]); ref[:, 3] = 1
q = ref.copy(); q[:, :3] += rng.normal(0, 0.003, (n, 3)).astype(np.float32); q = q[rng.permutation(n)]
dref, dq = torch.from_numpy(ref).cuda(), torch.from_numpy(q).cuda()
for occ in [float(v) for v in sys.argv[3].split(",")]:
    if occ > 0: os.environ["PCC_OCC"] = str(occ)
    s = GridSearch(0).setInputCloud(dref, k_hint=k); s.setTiming(True)
    best = 1e9
    for _ in range(3):
        s.nearestKSearch(dq, k); best = min(best, s.lastKernelMs())
    print(json.dumps(dict(n=n, k=k, occ=occ, grid=s.grid_info(), knn_ms=round(best, 3), gqps=round(n / best / 1e6, 3))), flush=True)
