#!/usr/bin/env python
"""Summarise the source page of an ncu report: stall-reason totals and the top-sampled SASS instructions.
    ncu -i rep.ncu-rep --page source --csv > src.csv ; python tools/ncu_src_top.py src.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
h0 = next(i for i, r in enumerate(rows) if "# Samples" in r)   # one or two title rows, depending on --print-source
hdr = rows[h0]
ix = {h: i for i, h in enumerate(hdr)}   # (duplicate "Source" columns: the last one, the SASS text, wins)
data = [r for r in rows[h0 + 1:] if len(r) == len(hdr)]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {h: 0 for h in stalls}
total = 0
for r in data:
    try:
        total += int(r[ix["# Samples"]])
    except ValueError:
        continue
    for h in stalls:
        tot[h] += int(r[ix[h]] or 0)
print("total samples", total)
for h, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]:
    print(f"  {h:28s} {v:8d} {v / max(total, 1):6.1%}")
order = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]] or 0))[:top]
for i in sorted(order):
    r = data[i]
    st = sorted(((int(r[ix[h]] or 0), h) for h in stalls), reverse=True)[:2]
    print(f"{i:5d} {int(r[ix['# Samples']]):7d} {r[ix['Instructions Executed']]:>10s}  {r[ix['Source']].strip()[:70]:70s} {st[0][1]}={st[0][0]} {st[1][1]}={st[1][0]}")
