import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pointcloudcomparator_b200 import synth
from pointcloudcomparator_b200.search import GridSearch
n = int(sys.argv[1])
src, tgt, T = synth.icp_pair(n, 4001, stride4=True)
ds, dt = torch.from_numpy(src).cuda(), torch.from_numpy(tgt).cuda()
for kh in [int(v) for v in sys.argv[2].split(",")]:
    s = GridSearch(0); s.setInputCloud(dt, k_hint=kh); torch.cuda.synchronize()
    tas = []
    if os.environ.get("PCC_ICP_TRACE"): s.setTiming(True)
    for rep in range(3):                       # the first call also pays for the scratch allocations
        t0 = time.perf_counter(); r = s.icpAlign(ds, 20); torch.cuda.synchronize(); tas.append(round((time.perf_counter() - t0) * 1e3, 1))
    print(json.dumps(dict(k_hint=kh, grid=s.grid_info(), align_ms=tas, it=r["iterations"], fitness=r["fitness"])), flush=True)
s = GridSearch(0); s.setInputCloud(dt, k_hint=17); s.setTiming(True)
for _ in range(3): s.meanNeighbourDistance(None, 16)
print("meandist16 self ms", s.lastKernelMs())
