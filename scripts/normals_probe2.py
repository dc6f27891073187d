"""Developer probe: where do GPU normals differ from the oracle's by more than 1e-5, and is that exactly the ill-conditioned set?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from pointcloudcomparator_b200 import synth
from pointcloudcomparator_b200.search import GridSearch

def gaps(ref, rows):
    g = np.zeros(len(rows)); 
    for i, nb in enumerate(rows):
        if len(nb) < 3: g[i] = np.nan; continue
        p = ref[nb].astype(np.float64); c = np.cov(p.T, bias=True); w = np.linalg.eigvalsh(c)
        g[i] = (w[1] - w[0]) / max(np.abs(c).max(), 1e-300)
    return g

for name, n, mode in (("knn50", 60000, "knn"), ("r0.03", 40000, "rad")):
    ref = synth.room(n, 1001)
    tree = oracle.KdTree(ref)
    if mode == "knn":
        s = GridSearch(0).setInputCloud(ref, k_hint=50); got = s.normalsKnn(None, 50); exp = oracle.normals_knn(ref, 50, tree=tree)
        rows = list(tree.knn(ref, 50)[0])
    else:
        s = GridSearch(0).setInputCloud(ref, cell_hint=0.03); got = s.normalsRadius(None, 0.03); exp = oracle.normals_radius(ref, 0.03, tree=tree)
        off, idx, _ = tree.radius(ref, 0.03); rows = [idx[off[i]:off[i + 1]] for i in range(n)]
    ok = ~np.isnan(exp[:, 0])
    err = np.linalg.norm(got[:, :3] - exp[:, :3], axis=1)
    g = gaps(ref, rows)
    bad = ok & (err > 1e-5)
    print(name, "n", ok.sum(), "err q50/q99/q999/max", np.quantile(err[ok], [.5, .99, .999, 1]), "frac>1e-5", bad.mean())
    print("  gap of bad: max", np.nanmax(g[bad]) if bad.any() else None, "q50", np.nanmedian(g[bad]) if bad.any() else None, " gap of all q01/q10/q50", np.nanquantile(g[ok], [.01, .1, .5]))
    print("  err*gap max", np.nanmax((err * g)[ok]), "q999", np.nanquantile((err * g)[ok], .999))
    for thr in (0.3, 0.1, 0.03, 0.01):
        w = ok & (g > thr); print(f"  gap>{thr}: frac {w.mean():.4f} max err {err[w].max():.3e}")
    cur = np.abs(got[ok, 3] - exp[ok, 3]) / np.maximum(np.abs(exp[ok, 3]), 1e-12)
    print("  curvature rel err q50/q99/max", np.quantile(cur, [.5, .99, 1]), "abs max", np.abs(got[ok, 3] - exp[ok, 3]).max())
