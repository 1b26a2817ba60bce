#!/usr/bin/env python3
"""Static SASS evidence per kernel of the built library: how often the opcodes that prove (or rule out) a hardware path occur.
usage: sass_markers.py [lib.so] > profiles/r2_sass_markers.txt        (needs cuobjdump + c++filt; no GPU)"""
import collections
import os
import re
import subprocess
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "minizip-ng_b200", "libmz_strm_cuda.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True).stdout
cols = [("UBLKCP", r"\bUBLKCP"), ("SYNCS", r"\bSYNCS"), ("LDG.128", r"\bLDG\.E[.\w]*\.128"), ("STG.128", r"\bSTG\.E[.\w]*\.128"), ("LDS.128", r"\bLDS\.128"),
        ("ATOMS", r"\bATOMS"), ("SHFL", r"\bSHFL"), ("VOTE", r"\bVOTE"), ("BAR", r"\bBAR\."), ("tensor", r"\b(HMMA|IMMA|UTC\w*MMA|QMMA|OMMA)")]
kern, counts, instrs, order = None, collections.defaultdict(collections.Counter), collections.Counter(), []
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
        kern = re.sub(r"^void |mzc::|\(.*$|\(int\)|\(bool\)", "", kern)
        order.append(kern)
        continue
    if kern and re.match(r"\s*/\*[0-9a-f]{4,}\*/", line):
        instrs[kern] += 1
        for name, rx in cols:
            if re.search(rx, line):
                counts[kern][name] += 1
print("profiles/r2_sass_markers.txt -- cuobjdump -sass %s (sm_100a), opcode occurrences per kernel (static counts; tools/sass_markers.py)" % os.path.relpath(lib, ROOT))
print("UBLKCP = cp.async.bulk (TMA bulk copy engine), SYNCS = mbarrier operations, LDG/STG .128 = 128-bit global loads/stores;")
print("tensor = HMMA/IMMA/UTC*MMA opcodes: none, as intended (there is no dense contraction on this path)")
print("%-40s" % "kernel" + "".join("%8s" % c for c, _ in cols) + "%8s" % "instrs")
for k in order:
    print("%-40s" % k[:40] + "".join("%8d" % counts[k][c] for c, _ in cols) + "%8d" % instrs[k])
