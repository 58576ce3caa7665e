#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list: count, total / mean us, share."""
import csv, sys
from collections import defaultdict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
agg = defaultdict(lambda: [0, 0.0])
data = rows[1:][skip:]
for r in data:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    us = v / 1000.0 if unit in ("ns", "nsecond") else v * 1000.0 if unit in ("ms", "msecond") else v
    name = r[ix["Kernel Name"]].split("(")[0][:60]
    agg[name][0] += 1; agg[name][1] += us
tot = sum(v[1] for v in agg.values())
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:62s} n={n:4d} total {us:10.1f} us  mean {us / n:9.2f} us  {100 * us / tot:5.1f} %")
print(f"total {tot:.1f} us")
