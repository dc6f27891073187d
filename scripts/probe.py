"""Developer probe: kNN kernel time vs grid occupancy target / k / cloud kind (not the bench contract).
usage: probe.py N kind k occ1,occ2,...   (occ 0 = library default)"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pointcloudcomparator_b200 import synth
from pointcloudcomparator_b200 import _lib
_lib.SO_PATH = os.path.abspath(os.environ.get("PCC_SO") or _lib.SO_PATH)          # developer builds from scripts/build_variant.sh
from pointcloudcomparator_b200.search import GridSearch

n = int(sys.argv[1]); kind = sys.argv[2]; ks = [int(v) for v in sys.argv[3].split(",")]; occs = [float(v) for v in sys.argv[4].split(",")]
cache = f"/tmp/pcc_probe_{kind}_{n}.npz"          # the clouds take ~20 s to generate; successive probes in one gpurun call share them
if os.path.exists(cache):
    z = np.load(cache); ref, q = z["ref"], z["q"]
else:
    ref = synth.room(n, 4001, size=(10, 10, 3), stride4=True) if kind == "surface" else synth.uniform(n, 5001, 10.0, stride4=True)
    q = synth.sweep_queries(ref, n, 5002, 0.01, stride4=True)
    np.savez(cache, ref=ref, q=q)
dref, dq = torch.from_numpy(ref).cuda(), torch.from_numpy(q).cuda()
for k in ks:
    for occ in occs:
        if occ > 0: os.environ["PCC_OCC"] = str(occ)
        else: os.environ.pop("PCC_OCC", None)
        s = GridSearch(0)
        s.setInputCloud(dref, k_hint=k); torch.cuda.synchronize()
        s.setTiming(True)
        best = 1e9
        for _ in range(4):
            out = s.nearestKSearch(dq, k); best = min(best, s.lastKernelMs())
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(5):
            out = s.nearestKSearch(dq, k)
        torch.cuda.synchronize(); call_ms = (time.perf_counter() - t0) / 5 * 1e3
        bself = 1e9
        for _ in range(2):
            md = s.meanNeighbourDistance(None, k); bself = min(bself, s.lastKernelMs())
        bq = 16 + 16 + 8 * k
        print(json.dumps(dict(kind=kind, n=n, k=k, occ=occ, grid=s.grid_info(), so=os.path.basename(_lib.SO_PATH), knn_ms=round(best, 3), call_ms=round(call_ms, 3), gqps=round(n / best / 1e6, 3),
                              frac=round(n * bq / (best * 1e-3) / 6533.8e9, 4), meandist_self_ms=round(bself, 3))), flush=True)
        del s, out, md
