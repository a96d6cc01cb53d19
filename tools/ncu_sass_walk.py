#!/usr/bin/env python
"""Walk the SASS of an ncu source page in address order: prints the barrier / TMEM / MMA instructions and every
instruction with >= MIN samples, with the samples accumulated since the previous printed line.
    ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass > src.csv ; python tools/ncu_sass_walk.py src.csv [MIN]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
mins = int(sys.argv[2]) if len(sys.argv) > 2 else 200
h0 = next(i for i, r in enumerate(rows) if "# Samples" in r)
hdr = rows[h0]
S = hdr.index("# Samples")


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


sass, cur = {}, None
for r in rows[h0 + 1:]:
    if len(r) <= S:
        continue
    if r[0] != "":
        cur = r[0]
        continue
    if r[2].startswith("0x"):
        sass[int(r[2], 16)] = (cur, r[3].strip(), num(r[S]))
addrs = sorted(sass)
base = addrs[0]
print(len(addrs), "instructions,", sum(v[2] for v in sass.values()), "samples")
acc = 0
KEYS = ("SYNCS", "LDTM", "STTM", "BAR", "VOTE", "UTCHMMA", "UTCBAR", "WARPSYNC", "EXIT", "UTMALDG", "FENCE", "ELECT", "BRA")
for a in addrs:
    line, txt, n = sass[a]
    acc += n
    if any(k in txt for k in KEYS) or n >= mins:
        print(f"{a - base:6x} L{line:>4s} {n:6d} acc={acc:6d} {txt[:90]}")
        acc = 0
