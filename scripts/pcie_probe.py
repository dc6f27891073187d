"""Developer probe: pinned-memory PCIe bandwidth of the box (D2H alone, H2D alone, both at once) -- the floor of bench.py's e2e."""
import time, torch
n = 1_280_000_000 // 4
d = torch.empty(n, dtype=torch.float32, device="cuda"); h = torch.empty(n, dtype=torch.float32).pin_memory()
d2 = torch.empty(n // 8, dtype=torch.float32, device="cuda"); h2 = torch.empty(n // 8, dtype=torch.float32).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=3):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
a = t(lambda: h.copy_(d, non_blocking=True)); print(f"D2H 1.28 GB: {a*1e3:.2f} ms = {1.28/a:.1f} GB/s")
b = t(lambda: d2.copy_(h2, non_blocking=True)); print(f"H2D 0.16 GB: {b*1e3:.2f} ms = {0.16/b:.1f} GB/s")
def both():
    with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
c = t(both); print(f"both at once: {c*1e3:.2f} ms = {1.44/c:.1f} GB/s total")
def chunks():
    k = 10
    for i in range(k):
        with torch.cuda.stream(s1 if i % 2 == 0 else s2): h[i * (n // k):(i + 1) * (n // k)].copy_(d[i * (n // k):(i + 1) * (n // k)], non_blocking=True)
e = t(chunks); print(f"D2H in 10 chunks on 2 streams: {e*1e3:.2f} ms = {1.28/e:.1f} GB/s")
