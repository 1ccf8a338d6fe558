#!/usr/bin/env python
"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) over the last `n` launches (one step)."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        h, start = r, i
        break
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
data = [(r[ki], float(r[vi].replace(",", ""))) for r in rows[start + 2:] if len(r) > vi]
agg = collections.OrderedDict()
for k, v in data[-n:]:
    k = k.split("(")[0][:70]
    agg.setdefault(k, [0, 0]); agg[k][0] += v; agg[k][1] += 1
tot = sum(v[0] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{v[0]/1e6:8.3f} ms x{v[1]:3d} {100*v[0]/tot:5.1f}%  {k}")
print(f"{tot/1e6:8.3f} ms total over the last {n} launches")
