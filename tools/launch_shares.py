#!/usr/bin/env python3
"""Per-kernel time shares from an ncu launch list (`--metrics gpu__time_duration.sum --csv`):
   python tools/launch_shares.py gpurun_out/launches_r2h.csv > profiles/r2h_launch_shares.md"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = rows[hi]
ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
acc = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= vi or not r[vi]:
        continue
    v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(r[ui], 1e-3)
    k = r[ki].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
    a = acc.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in acc.values())
print("| kernel | launches | total us | us per launch | share |\n|---|---|---|---|---|")
for k, (n, t) in acc.items():
    print("| %s | %d | %.1f | %.1f | %.1f %% |" % (k, n, t, t / n, 100 * t / tot))
