"""Runs the hot-path stage of each BASELINE.json config at its named size on one GPU and prints one JSON line per
stage (device-resident inputs, wall clock around the synchronous C-ABI call, best of 3; build_ms = first build on a fresh index, scratch
allocation included, build_steady_ms = the same build repeated on that index).  Not the bench contract:
this is the evidence for DESIGN.md's per-config table and a crash test at full size."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pointcloudcomparator_b200 import synth
from pointcloudcomparator_b200.search import GridSearch

def timed(fn, reps=3):
    best, out = 1e18, None
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3, out

def emit(**kw):
    print(json.dumps(kw), flush=True)

which = sys.argv[1].split(",") if len(sys.argv) > 1 else ["c1", "c2", "c3", "c4"]
if "c1" in which:   # default path on the 100k room: NormalEstimation k=50, RegionGrowing neighbour table k=100
    a = torch.from_numpy(synth.room(100_000, 1001, stride4=True)).cuda()
    s = GridSearch(0); tb, _ = timed(lambda: s.setInputCloud(a, k_hint=50), 1); tb2, _ = timed(lambda: s.setInputCloud(a, k_hint=50), 2)
    tn, n = timed(lambda: s.normalsKnn(None, 50))
    s2 = GridSearch(0); s2.setInputCloud(a, k_hint=100)
    tt, tab = timed(lambda: s2.nearestKSearch(None, 100))
    emit(config="C1 room-100k", build_ms=tb, build_steady_ms=tb2, normals_k50_ms=tn, table_k100_ms=tt, nan_normals=int(torch.isnan(n[:, 0]).sum()))
if "c2" in which:   # -n noise analysis on 1M points
    a = torch.from_numpy(synth.room(1_000_000, 2001, stride4=True)).cuda()
    for mk in (16, 50):
        s = GridSearch(0); tb, _ = timed(lambda: s.setInputCloud(a, k_hint=mk + 1), 1); tb2, _ = timed(lambda: s.setInputCloud(a, k_hint=mk + 1), 2)
        td, d = timed(lambda: s.meanNeighbourDistance(None, mk))
        tt, r = timed(lambda: s.sorThreshold(d, a.shape[0], 1.5))
        emit(config="C2 noise-1M", mean_k=mk, build_ms=tb, build_steady_ms=tb2, mean_dist_ms=td, threshold_ms=tt, kept=r["kept"], mean=r["mean"], stddev=r["stddev"])
if "c3" in which:   # -e Euclidean clusters on the 5M scene
    pts, ids = synth.scene(5_000_000, 3001, stride4=True)
    a = torch.from_numpy(pts).cuda()
    for tol in (0.02, 0.05):
        s = GridSearch(0); tb, _ = timed(lambda: s.setInputCloud(a, cell_hint=tol), 1); tb2, _ = timed(lambda: s.setInputCloud(a, cell_hint=tol), 2)
        tc, (lab, sizes) = timed(lambda: s.euclideanClusters(tol, 100, 250000))
        tr, (off, idx, d2) = timed(lambda: s.radiusSearch(None, tol), 1)
        emit(config="C3 scene-5M", tolerance=tol, build_ms=tb, build_steady_ms=tb2, clusters_ms=tc, n_clusters=int(sizes.numel()), largest=int(sizes[0]) if sizes.numel() else 0,
             radius_csr_ms=tr, mean_neighbours=float(off[-1]) / a.shape[0], grid=s.grid_info())
if "c4" in which:   # -i ICP 10M vs 10M
    src, tgt, T = synth.icp_pair(10_000_000, 4001, stride4=True)
    ds, dt = torch.from_numpy(src).cuda(), torch.from_numpy(tgt).cuda()
    s = GridSearch(0); tb, _ = timed(lambda: s.setInputCloud(dt, k_hint=32), 1); tb2, _ = timed(lambda: s.setInputCloud(dt, k_hint=32), 2)     # coarse cells: the first iterations search from far away
    t1, r1 = timed(lambda: s.icpStep(ds.clone(), None), 2)
    ta, r = timed(lambda: s.icpAlign(ds, 20), 1)
    emit(config="C4 icp-10M", build_ms=tb, build_steady_ms=tb2, one_pass_ms=t1, align_ms=ta, iterations=r["iterations"], converged=r["converged"], fitness=r["fitness"],
         T_err=float(np.abs(r["T"] - np.linalg.inv(T)).max()))
