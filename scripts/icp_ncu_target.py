"""Profiling target: the first ICP passes of the 10 M vs 10 M pair (C4), one pcc_icp_step per pass."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointcloudcomparator_b200 import synth
from pointcloudcomparator_b200.search import GridSearch, umeyama_from_sums
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 4
src, tgt, T = synth.icp_pair(n, 4001, stride4=True)
s = GridSearch(0).setInputCloud(torch.from_numpy(tgt).cuda(), k_hint=32)
cur = torch.from_numpy(src).cuda().clone()
Tstep = None
for it in range(passes):
    cnt, sums, _, _ = s.icpStep(cur, Tstep)
    Tstep = umeyama_from_sums(sums, cnt)
torch.cuda.synchronize()
print("done", s.grid_info(), cnt)
