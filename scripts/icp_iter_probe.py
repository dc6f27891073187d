import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pointcloudcomparator_b200 import synth
from pointcloudcomparator_b200.search import GridSearch, umeyama_from_sums
n = int(sys.argv[1])
src, tgt, T = synth.icp_pair(n, 4001, stride4=True)
dt = torch.from_numpy(tgt).cuda()
for kh in [int(v) for v in sys.argv[2].split(",")]:
    s = GridSearch(0); s.setInputCloud(dt, k_hint=kh); torch.cuda.synchronize()
    cur = torch.from_numpy(src).cuda().clone()
    Tstep = None; times = []; mses = []
    for it in range(12):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        cnt, sums, _, _ = s.icpStep(cur, Tstep)
        torch.cuda.synchronize(); times.append(round((time.perf_counter() - t0) * 1e3, 2))
        Tstep = umeyama_from_sums(sums, cnt); mses.append(float(sums[15] / cnt))
    print(json.dumps(dict(k_hint=kh, cell=s.grid_info()["cell"], ms_per_iter=times, rms_mm=[round(1e3 * m ** 0.5, 2) for m in mses])), flush=True)
