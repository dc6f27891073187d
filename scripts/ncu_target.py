"""Minimal profiling target: build the 10 M-point grid, run the k=16 kNN pass a few times (same workload as bench.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointcloudcomparator_b200 import synth
from pointcloudcomparator_b200.search import GridSearch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 16
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
mode = sys.argv[4] if len(sys.argv) > 4 else "knn"
ref = synth.room(n, 4001, size=(10.0, 10.0, 3.0), stride4=True)
qry = synth.sweep_queries(ref, n, seed=5002, sigma=0.01, stride4=True)
s = GridSearch(0).setInputCloud(torch.from_numpy(ref).cuda(), k_hint=k)
dq = torch.from_numpy(qry).cuda()
for _ in range(reps):
    if mode == "knn":
        out = s.nearestKSearch(dq, k)
    elif mode == "meandist":
        out = s.meanNeighbourDistance(None, k)
torch.cuda.synchronize()
print("done", s.grid_info())
