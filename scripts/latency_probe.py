"""Developer probe: per-call latency of pcc_knn for small device-resident batches (k = 16 and 50)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointcloudcomparator_b200 import synth
from pointcloudcomparator_b200.search import GridSearch
ref = torch.from_numpy(synth.room(1_000_000, 1001, stride4=True)).cuda()
for k in (16, 50):
    s = GridSearch(0).setInputCloud(ref, k_hint=k)
    for nq in (1, 512, 1023, 1024, 4096, 65536):
        q = ref[:nq].clone()
        for _ in range(5): s.nearestKSearch(q, k)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(200): s.nearestKSearch(q, k)
        torch.cuda.synchronize()
        print(json.dumps(dict(k=k, nq=nq, us_per_call=round((time.perf_counter() - t0) / 200 * 1e6, 1))), flush=True)
