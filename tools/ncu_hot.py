#!/usr/bin/env python
"""Summarise `ncu --page source --csv --print-source sass` output: stall-reason totals and the hottest SASS lines."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
rows = rows[starts[which]:starts[which + 1]]
print(rows[0][1][:100])
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("total samples", tot)
agg = {s: sum(int(r[ix[s]] or 0) for r in data) for s in stalls}
for s, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]:
    print(f"  {s:28s} {v:8d} {100.0*v/max(tot,1):5.1f}%")
print("hottest lines:")
order = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]] or 0))[:top]
for i in sorted(order):
    r = data[i]
    best = max(stalls, key=lambda s: int(r[ix[s]] or 0))
    print(f"  {i:5d} {int(r[ix['# Samples']]):6d} {best:18s} exec={r[ix['Instructions Executed']]:>9s}  {r[ix['Source']].strip()[:90]}")
