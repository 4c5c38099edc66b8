#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name.
    python profiles/launch_summary.py launches.csv [first_marker_kernel]
With a marker (e.g. prefetch_l2_kernel) only the launches between the last two marker launches are counted (= one step)."""
import csv
import sys
from collections import OrderedDict

rows = []
with open(sys.argv[1]) as fh:
    lines = [l for l in fh if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1e3 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1e3
        rows.append((r["Kernel Name"], us))
marker = sys.argv[2] if len(sys.argv) > 2 else None
if marker:
    idx = [i for i, (k, _) in enumerate(rows) if marker in k]
    if len(idx) >= 2:
        rows = rows[idx[-2]:idx[-1]]
agg = OrderedDict()
for k, us in rows:
    k = k.split("(")[0][:60]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1; a[1] += us
tot = sum(a[1] for a in agg.values())
print(f"# launches: {len(rows)}   total GPU time: {tot:.1f} us")
print(f"{'kernel':62s} {'launches':>8s} {'us':>10s} {'share':>7s}")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:62s} {n:8d} {us:10.1f} {100 * us / tot:6.1f}%")
