"""Developer probe: NormalEstimation kNN time vs cloud size (self query).  usage: normals_probe.py k n1,n2,..."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointcloudcomparator_b200 import synth
from pointcloudcomparator_b200.search import GridSearch
k = int(sys.argv[1])
for n in [int(v) for v in sys.argv[2].split(",")]:
    a = torch.from_numpy(synth.room(n, 1001, stride4=True)).cuda()
    s = GridSearch(0).setInputCloud(a, k_hint=k)
    best = 1e9
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter(); out = s.normalsKnn(None, k); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    print(json.dumps(dict(n=n, k=k, normals_ms=round(best * 1e3, 3), exact_only=bool(os.environ.get("PCC_EXACT_ONLY")))), flush=True)
