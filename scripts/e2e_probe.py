"""Developer probe: pcc_knn(PCC_HOST) end to end (pinned buffers) vs the pipeline chunk size (PCC_PIPE_CHUNK_LOG2)."""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pointcloudcomparator_b200 import synth, _lib
from pointcloudcomparator_b200.search import GridSearch
n = 10_000_000
cache = f"/tmp/pcc_probe_surface_{n}.npz"
if os.path.exists(cache):
    z = np.load(cache); ref, q = z["ref"], z["q"]
else:
    ref = synth.room(n, 4001, size=(10, 10, 3), stride4=True); q = synth.sweep_queries(ref, n, 5002, 0.01, stride4=True); np.savez(cache, ref=ref, q=q)
s = GridSearch(0).setInputCloud(torch.from_numpy(ref).cuda(), k_hint=16)
hq = torch.from_numpy(q).pin_memory().numpy()
hi = torch.empty((n, 16), dtype=torch.int32).pin_memory(); hd = torch.empty((n, 16), dtype=torch.float32).pin_memory()
L = _lib.lib(); keff = C.c_int()
def step(): _lib.check(L.pcc_knn(s._h, hq.ctypes.data, n, hq.strides[0], 16, hi.data_ptr(), hd.data_ptr(), C.byref(keff), _lib.HOST, None))
step(); step(); torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): step()
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print(f"chunk_log2={os.environ.get('PCC_PIPE_CHUNK_LOG2', '20')} e2e {dt*1e3:.2f} ms/step = {n/dt/1e9:.3f} G queries/s, {1.44/dt:.1f} GB/s PCIe")
