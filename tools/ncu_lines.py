#!/usr/bin/env python
"""Summarise an ncu `--page source --print-source cuda,sass --csv` dump per CUDA source line:
samples, warp instructions, average active threads.  usage: ncu_lines.py dump.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out, cur_file, hdr = [], None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if r[0] == "Function Name" or hdr is None:
        continue
    if r[0] != "" and r[0].isdigit():
        d = dict(zip(hdr[4:], r[4:]))

        def num(k):
            try:
                return float(d.get(k, "0") or 0)
            except ValueError:
                return 0.0
        out.append((cur_file, int(r[0]), r[1].strip()[:70], num("# Samples"), num("Instructions Executed"),
                    num("Thread Instructions Executed"), num("stall_long_sb"), num("stall_short_sb"), num("stall_lg"),
                    num("stall_wait"), num("stall_branch_resolving"), num("stall_mio")))
tot_s = sum(o[3] for o in out) or 1
tot_i = sum(o[4] for o in out) or 1
tot_t = sum(o[5] for o in out)
print("total samples %d, warp instr %.3e, thread instr %.3e, avg active threads %.2f" % (tot_s, tot_i, tot_t, tot_t / tot_i))
print("%-16s %5s %6s %6s %5s  %5s %5s %5s %5s %5s  %s" % ("file", "line", "smp%", "ins%", "thr", "longsb", "shrt", "lg", "wait", "mio", "source"))
for o in sorted(out, key=lambda o: -o[3])[:top]:
    thr = o[5] / o[4] if o[4] else 0
    print("%-16s %5d %6.2f %6.2f %5.1f  %5.0f %5.0f %5.0f %5.0f %5.0f  %s" % (o[0], o[1], 100 * o[3] / tot_s, 100 * o[4] / tot_i, thr,
                                                                  100 * o[6] / max(o[3], 1), 100 * o[7] / max(o[3], 1), 100 * o[8] / max(o[3], 1),
                                                                  100 * o[9] / max(o[3], 1), 100 * o[11] / max(o[3], 1), o[2]))
