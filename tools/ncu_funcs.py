#!/usr/bin/env python
"""Aggregate an ncu cuda,sass source dump by function of sf_core.cuh: usage ncu_funcs.py dump.csv core.cuh"""
import csv
import re
import sys

funcs = []
for i, l in enumerate(open(sys.argv[2]), 1):
    if l.startswith("SF_FN") or l.startswith("template"):
        m = re.search(r"(\w+)\(", l)
        if m:
            funcs.append((i, m.group(1)))
rows = list(csv.reader(open(sys.argv[1])))
cur = None
hdr = None
agg = {}


def num(x):
    try:
        return float(x)
    except ValueError:
        return 0.0


for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if r[0] == "Function Name" or hdr is None:
        continue
    if r[0].isdigit() and len(r) > 8:
        d = dict(zip(hdr[4:], r[4:]))
        ln = int(r[0])
        fn = cur
        if cur == "sf_core.cuh":
            c = [f for l, f in funcs if l <= ln]
            fn = c[-1] if c else "head"
        a = agg.setdefault(fn, [0, 0, 0])
        a[0] += num(d.get("# Samples", 0))
        a[1] += num(d.get("Instructions Executed", 0))
        a[2] += num(d.get("Thread Instructions Executed", 0))
tot = sum(a[0] for a in agg.values()) or 1
toti = sum(a[1] for a in agg.values()) or 1
for fn, a in sorted(agg.items(), key=lambda x: -x[1][0])[:28]:
    print("%-22s samples %5.1f%%  instr %5.1f%%  thr %.1f" % (fn, 100 * a[0] / tot, 100 * a[1] / toti, a[2] / max(a[1], 1)))
