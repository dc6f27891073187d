"""Developer probe: time the kNN kernel at several sizes / k / occupancy targets (not the bench contract)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pointcloudcomparator_b200 import synth
from pointcloudcomparator_b200.search import GridSearch

def run(kind, n, nq, k, occ=None, reps=3):
    if occ: os.environ["PCC_OCC"] = str(occ)
    else: os.environ.pop("PCC_OCC", None)
    ref = synth.room(n, 4001, size=(10, 10, 3), stride4=True) if kind == "surface" else synth.uniform(n, 5001, 10.0, stride4=True)
    q = synth.sweep_queries(ref, nq, 5002, 0.01, stride4=True)
    dref, dq = torch.from_numpy(ref).cuda(), torch.from_numpy(q).cuda()
    s = GridSearch(0)
    t0 = time.time(); s.setInputCloud(dref, k_hint=k); torch.cuda.synchronize(); tb = time.time() - t0
    s.setTiming(True)
    best = 1e9; tot = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.time()
        idx, d2, _ = s.nearestKSearch(dq, k)
        torch.cuda.synchronize(); tot = min(tot, time.time() - t0)
        best = min(best, s.lastKernelMs())
    s.setTiming(True)
    bself = 1e9
    for _ in range(reps):
        md = s.meanNeighbourDistance(None, k); bself = min(bself, s.lastKernelMs())
    bytes_q = 16 * n / nq + 16 + 8 * k
    print(json.dumps(dict(kind=kind, n=n, nq=nq, k=k, occ=occ, grid=s.grid_info(), build_s=round(tb, 4), knn_kernel_ms=round(best, 3), knn_call_ms=round(tot * 1e3, 3),
                          gqps=round(nq / best / 1e6, 3), roofline_frac=round(nq * bytes_q / (best * 1e-3) / 6533.8e9, 4), meandist_self_ms=round(bself, 3))), flush=True)

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    for kind in ("surface", "uniform"):
        for occ in (None, 3.0, 8.0):
            run(kind, n, n, 16, occ)
    run("surface", n, n, 1); run("surface", n, n, 32); run("surface", n, n, 50, reps=1)
