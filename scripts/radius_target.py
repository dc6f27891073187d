import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointcloudcomparator_b200 import synth
from pointcloudcomparator_b200.search import GridSearch
n = int(sys.argv[1]); r = float(sys.argv[2])
pts, _ = synth.scene(n, 3001, stride4=True)
a = torch.from_numpy(pts).cuda()
s = GridSearch(0).setInputCloud(a, cell_hint=r)
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    off, idx, d2 = s.radiusSearch(None, r)
    torch.cuda.synchronize(); print("radius csr ms", (time.perf_counter() - t0) * 1e3, "pairs", int(off[-1]), flush=True)
    del off, idx, d2
