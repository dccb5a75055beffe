#!/usr/bin/env python
"""Stall-reason shares of an .ncu-rep (first kernel): usage ncu_stalls.py file.ncu-rep"""
import csv
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
d = dict(zip(rows[0], rows[2]))
items = []
for k, v in d.items():
    if "pcsamp_warps_issue_stalled" in k and not k.endswith("not_issued"):
        try:
            items.append((float(v), k.replace("smsp__pcsamp_warps_issue_stalled_", "")))
        except ValueError:
            pass
items.sort(reverse=True)
tot = sum(x for x, _ in items) or 1.0
for x, k in items[:12]:
    print("  %5.1f%%  %s" % (100 * x / tot, k))
