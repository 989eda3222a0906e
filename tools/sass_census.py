#!/usr/bin/env python
"""SASS census of libpss.so: per kernel, how many instructions of the classes that tell a Blackwell-era kernel
from a recompiled older one (B200_PROFILING.md "What proves a Blackwell-native kernel").
usage: python tools/sass_census.py [path/to/libpss.so] > profiles/<round>_sass_census.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "pyspecsdr_b200", "libpss.so")
CLASSES = ["DMMA", "DFMA", "HMMA", "UTCHMMA", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "LDGSTS", "UCGABAR", "CCTL",
           "LDG", "STG", "LDS", "STS", "ATOMS", "SHFL", "MUFU", "BAR", "F2F"]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, counts, totals = None, collections.OrderedDict(), collections.Counter()
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern).replace("void ", "")
        counts[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        op = m.group(1)
        counts[kern]["_all"] += 1
        for c in CLASSES:
            if op == c or op.startswith(c + "_") or (c in ("UCGABAR", "SYNCS", "UBLKCP", "UTMALDG", "UTMASTG") and op.startswith(c)):
                counts[kern][c] += 1
                totals[c] += 1
print(f"# SASS census of {os.path.relpath(lib, ROOT)} ({len(counts)} kernels, sm_100a)")
print("# classes: " + " ".join(CLASSES))
print(f"{'kernel':70s} {'insts':>7s} " + " ".join(f"{c:>7s}" for c in CLASSES))
for k, c in counts.items():
    print(f"{k[:70]:70s} {c['_all']:7d} " + " ".join(f"{c[x]:7d}" for x in CLASSES))
print(f"{'TOTAL':70s} {sum(c['_all'] for c in counts.values()):7d} " + " ".join(f"{totals[x]:7d}" for x in CLASSES))
