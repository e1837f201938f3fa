"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[h]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
d = defaultdict(list)
for r in rows[h + 1:]:
    if len(r) > vi:
        try:
            d[r[ki].split("(")[0][-60:]].append(float(r[vi]))
        except ValueError:
            pass
tot = sum(sum(v) / len(v) for v in d.values())
for k, v in d.items():
    avg = sum(v) / len(v) / 1000
    print(f"{k:62s} n={len(v):3d} avg={avg:9.2f} us  share={avg * 1000 / tot:6.1%}")
print(f"sum of per-kernel averages: {tot / 1000:.1f} us")
