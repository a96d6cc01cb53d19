#!/usr/bin/env python
"""Stall samples per CUDA source line (and, with --sass LINE, per SASS instruction of one line) from
    ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass > src.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h0 = next(i for i, r in enumerate(rows) if "# Samples" in r)
hdr = rows[h0]
S = hdr.index("# Samples")
stall_ix = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


data = [r for r in rows[h0 + 1:] if len(r) > S]
lines = [r for r in data if r[0] != ""]
total = sum(num(r[S]) for r in lines)
print("total samples", total)
top = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 40
for r in sorted(lines, key=lambda r: -num(r[S]))[:top]:
    st = sorted(((num(r[i]), h) for i, h in stall_ix), reverse=True)[:3]
    print(f"{r[0]:>5s} {num(r[S]):7d} {num(r[S]) / total:6.1%}  {r[1].strip()[:80]:80s} " + " ".join(f"{h[6:]}={v}" for v, h in st if v))
if "--sass" in sys.argv:
    want = sys.argv[sys.argv.index("--sass") + 1]
    on = False
    for r in data:
        if r[0] != "":
            on = r[0] == want
            continue
        if on:
            st = sorted(((num(r[i]), h) for i, h in stall_ix), reverse=True)[:2]
            print(f"   {num(r[S]):6d} {r[3][:70]:70s} " + " ".join(f"{h[6:]}={v}" for v, h in st if v))
