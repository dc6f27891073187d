"""BASELINE configs[4]: raw kNN / radius sweep on one GPU.
  kNN    : Q in {1 M, 10 M, 100 M} x k in {1, 2, 4, 8, 16, 32}, 10 M reference points, surface + uniform clouds
  radius : r in {1, 2, 5 cm} at Q = N = 10 M (surface), CSR count + fill (sorted rows)
Every line carries the HBM-roofline fraction of its algorithmic bytes (kNN: 16 N/Q + 16 + 8k per query; radius:
16 N/Q + 24 + 8 m_mean) against MEASURED_PEAKS.json.  Queries are reference points + N(0, 1 cm) drawn ON THE DEVICE (torch, seed 5002) --
the same distribution as synth.sweep_queries without the minutes of host generation a 100 M-query batch costs.
usage: python scripts/sweep_c5.py [out.jsonl] [--quick]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pointcloudcomparator_b200 import synth
from pointcloudcomparator_b200.search import GridSearch

out_path = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else "gpurun_out/sweep_c5.jsonl"
quick = "--quick" in sys.argv
N = 10_000_000
peak = 6533.8
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
f = open(out_path, "w")


def emit(d):
    f.write(json.dumps(d) + "\n"); f.flush(); print(json.dumps(d), flush=True)


def device_queries(dref, nq, seed):
    g = torch.Generator(device="cuda"); g.manual_seed(seed)
    pick = torch.randint(0, dref.shape[0], (nq,), device="cuda", generator=g)
    q = dref[pick].clone()
    q[:, :3] += torch.randn((nq, 3), device="cuda", generator=g) * 0.01
    return q


for kind in ("surface", "uniform"):
    ref = synth.room(N, 4001, size=(10.0, 10.0, 3.0), stride4=True) if kind == "surface" else synth.uniform(N, 5001, 10.0, stride4=True)
    dref = torch.from_numpy(ref).cuda()
    for Q in ((1_000_000, 10_000_000) if quick else (1_000_000, 10_000_000, 100_000_000)):
        dq = device_queries(dref, Q, 5002)
        for k in (1, 2, 4, 8, 16, 32):
            s = GridSearch(0).setInputCloud(dref, k_hint=k)
            s.setTiming(True)
            best, call = 1e9, 1e9
            for _ in range(3):
                torch.cuda.synchronize(); t0 = time.perf_counter()
                o = s.nearestKSearch(dq, k)
                torch.cuda.synchronize(); call = min(call, (time.perf_counter() - t0) * 1e3)
                best = min(best, s.lastKernelMs())
                del o
            b = 16.0 * N / Q + 16 + 8 * k
            emit(dict(op="knn", cloud=kind, n_ref=N, n_query=Q, k=k, kernel_ms=round(best, 3), call_ms=round(call, 3), gqps_kernel=round(Q / best / 1e6, 3), gqps_call=round(Q / call / 1e6, 3),
                      bytes_per_query=b, frac_kernel=round(Q * b / (best * 1e-3) / 1e9 / peak, 4), frac_call=round(Q * b / (call * 1e-3) / 1e9 / peak, 4), grid=s.grid_info()))
            del s
        if kind == "surface" and Q == 10_000_000:
            for r in (0.01, 0.02, 0.05):
                s = GridSearch(0).setInputCloud(dref, cell_hint=r)
                best = 1e9
                for _ in range(2):
                    torch.cuda.synchronize(); t0 = time.perf_counter()
                    off, idx, d2 = s.radiusSearch(dq, r)
                    torch.cuda.synchronize(); best = min(best, (time.perf_counter() - t0) * 1e3)
                    total = int(off[-1]); del idx, d2
                m = total / Q
                b = 16.0 * N / Q + 24 + 8 * m
                emit(dict(op="radius", cloud=kind, n_ref=N, n_query=Q, radius=r, pairs=total, mean_neighbours=round(m, 2), call_ms=round(best, 3), gqps_call=round(Q / best / 1e6, 3),
                          bytes_per_query=round(b, 1), frac_call=round(Q * b / (best * 1e-3) / 1e9 / peak, 4), grid=s.grid_info()))
                del s, off
        del dq
    del dref
