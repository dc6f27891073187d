"""Developer helper: per-source-line share of stall samples and executed instructions from an ncu report.
usage: python scripts/ncu_lines.py report.ncu-rep [top]"""
import csv, subprocess, sys
rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, r = rows[0], rows[2]
for n in ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
          "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "launch__registers_per_thread"]:
    if n in h: print(n, rows[1][h.index(n)], r[h.index(n)])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
fname, inst, smp = None, {}, {}
for row in csv.reader(src.splitlines()):
    if len(row) == 2 and row[0] == "File Name": fname = row[1].split("/")[-1]; continue
    if len(row) < 10 or row[0] == "Line No": continue
    if row[0] != "" and row[2] == "-":
        try:
            key = (fname, int(row[0]), row[1][:100]); inst[key] = inst.get(key, 0) + int(row[7]); smp[key] = smp.get(key, 0) + int(row[6])
        except ValueError: pass
ti, ts = sum(inst.values()) or 1, sum(smp.values()) or 1
print("total warp-instructions", ti, "samples", ts)
for k, v in sorted(smp.items(), key=lambda x: -x[1])[:top]: print(f"samples {v / ts * 100:5.1f}%  inst {inst[k] / ti * 100:5.1f}%  {k[0]}:{k[1]}  {k[2]}")
