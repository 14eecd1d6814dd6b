#!/usr/bin/env python
"""SASS census of toplingdb_b200/libb200c.so: per kernel, how many TMA / mbarrier / legacy async-copy instructions it holds
(cuobjdump -sass).  UBLKCP = cp.async.bulk (TMA bulk copy), SYNCS = mbarrier operations, UTMACMDFLUSH = bulk-group commit / wait,
LDGSTS = cp.async (Ampere-style).  usage: tools/sass_census.py > profiles/r02_sass_census.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "toplingdb_b200", "libb200c.so")], capture_output=True, text=True).stdout
kern, counts, total = None, collections.defaultdict(collections.Counter), collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0].replace("void ", "")
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1)
        total[kern] += 1
        for tag in ("UBLKCP", "UTMALDG", "UTMASTG", "UTMACMDFLUSH", "SYNCS", "LDGSTS", "UTMACCTL"):
            if op.startswith(tag):
                counts[kern][op] += 1
print("SASS census of toplingdb_b200/libb200c.so (cuobjdump -sass, sm_100a)\n")
print("%-58s %8s  %s" % ("kernel", "instrs", "TMA / mbarrier / cp.async instructions"))
for k in sorted(total, key=lambda k: -total[k]):
    c = counts.get(k)
    print("%-58s %8d  %s" % (k[:58], total[k], ", ".join(f"{op} x{n}" for op, n in sorted(c.items())) if c else "-"))
