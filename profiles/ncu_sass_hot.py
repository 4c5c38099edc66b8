#!/usr/bin/env python3
"""Hot SASS instructions of one kernel from `ncu -i X.ncu-rep --page source --csv [--launch-skip k --launch-count 1] > src.csv`:
    python profiles/ncu_sass_hot.py src.csv [min_share_percent]
prints, in address order, every instruction holding >= the share of the warp-stall samples plus every barrier / bulk copy / exit
(so that the samples between two barriers can be read as the cost of that phase)."""
import csv
import sys

lines = open(sys.argv[1]).read().splitlines()
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.6
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rd = csv.reader(lines[start:])
hdr = next(rd)
isrc, isamp, iex = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")


def num(v):
    try:
        return int(float(v))
    except ValueError:
        return 0


rows = [r for r in rd if len(r) > isamp and r[isamp] != hdr[isamp]]
tot = sum(num(r[isamp]) for r in rows) or 1
print(f"# {len(rows)} instructions, {tot} samples")
acc = 0
for i, r in enumerate(rows):
    s = num(r[isamp]); acc += s
    src = r[isrc]
    if s >= thr / 100 * tot or any(k in src for k in ("BAR.SYNC", "SYNCS", "UBLKCP", "ACQBULK", "EXIT", "UTCHMMA", "UTCBAR")):
        print(f"{i:5d} {s:6d} {100 * s / tot:5.1f}%  cum {100 * acc / tot:5.1f}%  ex {r[iex]:>7s}  {src[:110]}")
