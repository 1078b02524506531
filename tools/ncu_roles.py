#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` export: stall samples per instruction above a threshold, so the
waiting role of a warp-specialised kernel can be read off.  usage: ncu_roles.py file.csv [min_pct]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr = rows[1]
data = [r for r in rows[2:] if len(r) > 10]
iS, iSrc, iE = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


tot = sum(num(r[iS]) for r in data)
print("instructions", len(data), "samples", tot)
for n, r in enumerate(data):
    s = num(r[iS])
    if s > tot * thr / 100:
        print(f"{n:5d} {s:6d} {100 * s / tot:5.1f}%  x{r[iE]:>9s}  {r[iSrc].strip()[:100]}")
