"""Summarise an ncu launch list (gpu__time_duration.sum CSV): per-kernel mean of the last `reps` launches."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
tail = int(sys.argv[2]) if len(sys.argv) > 2 else 14
for r in rows[-tail:]:
    print(f"{float(r[vi].replace(',', '')) / 1e3:10.1f} us  {r[ki][:90]}")
