"""Compare the per-thread kNN kernel with the warp-owns-a-cell TMA variant at several query/reference ratios and check parity."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle
from pointcloudcomparator_b200 import synth
from pointcloudcomparator_b200.search import GridSearch
n = int(sys.argv[1]); ratios = [int(v) for v in sys.argv[2].split(",")]; mode = sys.argv[3]   # mode = value of PCC_CELL_KERNEL for this process
ref = synth.room(n, 4001, size=(10, 10, 3), stride4=True)
dref = torch.from_numpy(ref).cuda()
s = GridSearch(0).setInputCloud(dref, k_hint=16); s.setTiming(True)
tree = None
for ratio in ratios:
    q = synth.sweep_queries(ref, n * ratio, 5002, 0.01, stride4=True)
    dq = torch.from_numpy(q).cuda()
    best = 1e9
    for _ in range(3):
        idx, d2, _ = s.nearestKSearch(dq, 16); best = min(best, s.lastKernelMs())
    ok = None
    if ratio == ratios[0]:
        tree = tree or oracle.KdTree(ref)
        sel = np.random.default_rng(0).choice(q.shape[0], 20000, replace=False)
        oi, od, _ = tree.knn(q[sel], 16)
        gi = idx[torch.from_numpy(sel).cuda()].cpu().numpy(); gd = d2[torch.from_numpy(sel).cuda()].cpu().numpy()
        ok = bool(np.array_equal(gi, oi) and np.array_equal(gd.view(np.uint32), od.view(np.uint32)))
    print(json.dumps(dict(cell_kernel=mode, n=n, q=n * ratio, ms=round(best, 3), gqps=round(n * ratio / best / 1e6, 3), parity=ok)), flush=True)
    del dq, idx, d2
