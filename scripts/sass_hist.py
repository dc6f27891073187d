"""Summarise an `ncu --page source --csv` dump: instructions per warp by SASS region and opcode."""
import csv, sys
path, nwarps, step = sys.argv[1], float(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 80
rows = list(csv.reader(open(path)))
# keep the section of the main kernel (largest section)
secs, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name": cur = {"name": r[1], "rows": []}; secs.append(cur)
    elif cur is not None: cur["rows"].append(r)
sec = max(secs, key=lambda s: len(s["rows"]))
hdr = sec["rows"][0]; data = [r for r in sec["rows"][1:] if len(r) == len(hdr)]
ci, ti, src, smp = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
tot = sum(int(r[ci]) for r in data); tsm = sum(int(r[smp]) for r in data)
print(sec["name"][:80], "total warp inst", tot, "per warp", round(tot / nwarps, 1), "sass lines", len(data))
def opc(x):
    s = x[src].split()
    return (s[1] if s[0].startswith('@') else s[0]).split('.')[0]
for i in range(0, len(data), step):
    ch = data[i:i + step]
    n = sum(int(x[ci]) for x in ch); t = sum(int(x[ti]) for x in ch); sm = sum(int(x[smp]) for x in ch)
    ops = {}
    for x in ch: ops[opc(x)] = ops.get(opc(x), 0) + int(x[ci])
    top = sorted(ops.items(), key=lambda kv: -kv[1])[:5]
    print(f"sass[{i:5d}..] {n / nwarps:8.1f}/warp {100 * n / tot:5.1f}% thr {t / max(n, 1):5.1f} samples {100 * sm / max(tsm, 1):5.1f}%  " + ", ".join(f"{k}:{v / nwarps:.0f}" for k, v in top))
ops = {}
for x in data: ops[opc(x)] = ops.get(opc(x), 0) + int(x[ci])
print("by opcode:", ", ".join(f"{k}:{v / nwarps:.0f}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:14]))
names = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("stalls:", ", ".join(f"{n[6:]}:{sum(int(r[hdr.index(n)]) for r in data)}" for n in names if sum(int(r[hdr.index(n)]) for r in data) > 0))
