"""Profiling target: the C5 radius rows -- 10 M queries (points + 1 cm noise) vs the 10 M-point surface cloud, CSR result with sorted rows."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointcloudcomparator_b200 import synth
from pointcloudcomparator_b200.search import GridSearch
n = int(sys.argv[1]); r = float(sys.argv[2]); reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
ref = synth.room(n, 4001, size=(10.0, 10.0, 3.0), stride4=True)
qry = synth.sweep_queries(ref, n, seed=5002, sigma=0.01, stride4=True)
s = GridSearch(0).setInputCloud(torch.from_numpy(ref).cuda(), cell_hint=r)
dq = torch.from_numpy(qry).cuda()
for _ in range(reps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    off, idx, d2 = s.radiusSearch(dq, r)
    torch.cuda.synchronize(); print("radius csr ms", (time.perf_counter() - t0) * 1e3, "pairs", int(off[-1]), flush=True)
    del off, idx, d2
