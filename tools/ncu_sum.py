#!/usr/bin/env python
"""Sum gpu__time_duration per kernel name over the LAST `--last N` launches of an ncu --csv launch list."""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
last = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = list(csv.reader(open(path, errors="ignore")))
hdr = None
recs = []
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    if d["Metric Name"].startswith("gpu__time_duration"):
        v = float(d["Metric Value"].replace(",", ""))
        if d["Metric Unit"] in ("us", "usecond"):
            v *= 1e3
        elif d["Metric Unit"] in ("ms", "msecond"):
            v *= 1e6
        recs.append((d["Kernel Name"].split("(")[0][-60:], v))
if last:
    recs = recs[-last:]
agg = defaultdict(lambda: [0, 0.0])
for k, v in recs:
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v[1] for v in agg.values())
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:62s} n={n:5d} {t / 1e6:9.3f} ms {t / tot:6.1%}")
print(f"total {tot / 1e6:.3f} ms over {len(recs)} launches")
