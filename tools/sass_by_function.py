#!/usr/bin/env python
"""Static size of a kernel by source function: counts the SASS instructions of one kernel in
`nvdisasm --print-line-info` output per (file, enclosing function of the line).
usage: nvdisasm --print-line-info X.cubin | sass_by_function.py <kernel substring> <source file>..."""
import re
import sys

kern = sys.argv[1]
srcs = sys.argv[2:]
# function start lines per source file
starts = {}
for path in srcs:
    lst = []
    for n, line in enumerate(open(path), 1):
        m = re.match(r"^(?:SF_FN|SF_MFN|SF_COLD|__global__|__device__|template|static|inline)[^;]*?\b(\w+)\s*\(", line)
        if m and not line.strip().endswith(";"):
            lst.append((n, m.group(1)))
    starts[path.split("/")[-1]] = lst


def func_of(fname, line):
    best = "?"
    for n, name in starts.get(fname, []):
        if n <= line:
            best = name
        else:
            break
    return best


on, cur, counts, total = False, ("?", 0), {}, 0
for line in sys.stdin:
    if line.startswith(".text."):
        on = kern in line
        continue
    if not on:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        key = func_of(*cur) if cur[0] in starts else cur[0]
        counts[key] = counts.get(key, 0) + 1
        total += 1
print("total", total)
for k, v in sorted(counts.items(), key=lambda kv: -kv[1])[:40]:
    print("%6d %5.1f%%  %s" % (v, 100.0 * v / total, k))
