"""Summarises an `ncu --metrics gpu__time_duration.sum --csv --log-file X` launch list per kernel."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    if r[mi] != "gpu__time_duration.sum":
        continue
    v = float(r[vi].replace(",", ""))
    us = v / 1e3 if r[ui] in ("ns", "nsecond") else v if r[ui] in ("us", "usecond") else v * 1e3
    a = agg.setdefault(r[ki], [0, 0.0])
    a[0] += 1; a[1] += us
tot = sum(a[1] for a in agg.values())
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:72]:72s} n={n:4d} total_us={us:10.1f} share={us / tot:.3f} avg_us={us / n:.1f}")
