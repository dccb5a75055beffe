#!/usr/bin/env python
"""Instruction mix per kernel of a built library: usage sass_summary.py lib.so  (reads `cuobjdump -sass`).
Lists, per kernel, the SASS instruction count and the mnemonics that say what the kernel is made of
(shared / global / local memory, barriers, warp votes and reductions, bulk-copy and tensor-core ops)."""
import collections
import re
import subprocess
import sys

KEEP = ("LDS", "STS", "LDG", "STG", "LDL", "STL", "LDC", "BAR", "REDUX", "ATOMS", "ATOMG", "RED", "SHFL", "VOTE", "MATCH",
        "HMMA", "IMMA", "UTCHMMA", "UTCMMA", "UBLKCP", "UTMALDG", "UTMASTG", "CALL", "IMAD", "LOP3", "ISETP", "BRA", "WARPSYNC")
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
name, total, mix = None, collections.Counter(), collections.defaultdict(collections.Counter)
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and name:
        total[name] += 1
        mix[name][m.group(1)] += 1
for k in sorted(total, key=lambda n: -total[n]):
    print("%6d  %s" % (total[k], k))
    print("        " + "  ".join("%s %d" % (m, mix[k][m]) for m in KEEP if mix[k][m]))
tensor = sum(mix[k][m] for k in mix for m in ("HMMA", "IMMA", "UTCHMMA", "UTCMMA"))
print("tensor-core instructions in the library: %d (there is no dense contraction on this path)" % tensor)
