"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: python tools/launch_summary.py launches.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr, body = rows[hi], rows[hi + 1:]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in body:
    if len(r) <= mv:
        continue
    name = r[kn].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
    v = float(r[mv].replace(",", "")) / {"ns": 1e3, "us": 1.0, "usecond": 1.0, "ms": 1e-3}.get(r[mu], 1e3)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"{sum(a[0] for a in agg.values())} launches, {tot:.0f} us (cold-cache, serialised: compare shares)")
for k, a in sorted(agg.items(), key=lambda t: -t[1][1]):
    print(f"{k[:72]:72s} n={a[0]:4d} total={a[1]:10.1f} us share={100 * a[1] / tot:5.1f}%")
