import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["PCC_STATS"] = "1"
import torch
from pointcloudcomparator_b200 import synth
from pointcloudcomparator_b200 import _lib
_lib.SO_PATH = os.path.abspath(os.environ.get("PCC_SO", _lib.SO_PATH))          # developer builds from scripts/build_variant.sh
from pointcloudcomparator_b200.search import GridSearch
n = int(sys.argv[1]); kind = sys.argv[2]
ref = synth.room(n, 4001, size=(10, 10, 3), stride4=True) if kind == "surface" else synth.uniform(n, 5001, 10.0, stride4=True)
q = synth.sweep_queries(ref, n, 5002, 0.01, stride4=True)
dref, dq = torch.from_numpy(ref).cuda(), torch.from_numpy(q).cuda()
for occ in [float(v) for v in sys.argv[3].split(",")]:
    os.environ["PCC_OCC"] = str(occ)
    k = int(sys.argv[4]) if len(sys.argv) > 4 else 16
    s = GridSearch(0).setInputCloud(dref, k_hint=k); s.setTiming(True)
    s.nearestKSearch(dq, k); s.nearestKSearch(dq, k)
    print("occ", occ, s.grid_info(), "ms", s.lastKernelMs(), flush=True)
