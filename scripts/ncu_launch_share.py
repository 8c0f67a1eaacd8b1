#!/usr/bin/env python
"""Per-kernel share of a `ncu --metrics gpu__time_duration.sum --csv` launch list (profiles/*_ncu_launches.csv)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
MIN_US = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0  # only launches at least this long (the bench's 8-frame launches: 1000)
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[h]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = {}
for r in rows[h + 1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1e-3)
    if v < MIN_US:
        continue
    n = r[ki].split("(")[0].replace("void ", "")
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(v[1] for v in agg.values())
print("| kernel | launches | total us | mean us | share |\n|---|---|---|---|---|")
for n, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.1f | %.1f | %.3f |" % (n, v[0], v[1], v[1] / v[0], v[1] / tot))
