#!/usr/bin/env python
"""Per CUDA source line: global-memory sectors requested (ncu source page, "L2 Theoretical Sectors
Global") split into loads and stores.  usage: ncu_mem_lines.py dump.csv [top]
The dump is `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv`; the SASS rows carry the
access operation, the CUDA rows the totals per line."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
sass_view = False
lines = {}
cur_file = None
tot = {"Load": 0.0, "Store": 0.0}
# the cuda view lists, per source line, the totals; ops are only known per SASS instruction, so use the
# sass view rows ("Address" column filled) and map them back through the per-line view when present
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    d = dict(zip(hdr[4:], r[4:]))
    try:
        sec = float(d.get("L2 Theoretical Sectors Global") or 0)
        req = float(d.get("L1 Tag Requests Global") or 0)
    except ValueError:
        continue
    if sec == 0 or not r[0].isdigit():
        continue
    key = (cur_file, int(r[0]), r[1].strip()[:80])
    e = lines.setdefault(key, [0.0, 0.0, 0.0])
    e[0] += sec
    e[1] += req
    e[2] += float(d.get("L2 Theoretical Sectors Global Ideal") or 0)
total = sum(v[0] for v in lines.values()) or 1
print("total sectors %.3e, tag requests %.3e" % (total, sum(v[1] for v in lines.values())))
print("%-14s %5s %7s %7s %9s %6s  %s" % ("file", "line", "sect%", "cum%", "sectors", "s/req", "source"))
cum = 0
for k, v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
    cum += v[0]
    print("%-14s %5d %7.2f %7.2f %9.3e %6.1f  %s" % (k[0], k[1], 100 * v[0] / total, 100 * cum / total, v[0], v[0] / max(v[1], 1), k[2]))
