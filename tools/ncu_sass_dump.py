#!/usr/bin/env python3
"""Dump the SASS of an ncu source page (csv made with --page source --csv --print-source cuda,sass: grouped by CUDA line) back
in address order, with per-instruction executed counts (% of the kernel's warp-instructions), average active lanes, stall
samples and the CUDA line it belongs to, split at barriers like tools/ncu_segments.py.
usage: ncu_sass_dump.py <source.csv> [segment ...]  -> text on stdout"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
want = set(int(a) for a in sys.argv[2:])
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
ins = []
line = ""
for r in rows[hi + 1:]:
    if len(r) < 10:
        continue
    if r[0]:
        line = r[0]
    if r[2].startswith("0x"):
        ins.append((int(r[2], 16), r[3].strip(), int(r[7]), int(r[8]), int(r[6]), line))
ins.sort()
tot = sum(i[2] for i in ins)
seg = 0
for a, op, inst, thr, smp, line in ins:
    if not want or seg in want:
        print("%3d %5x %6.3f%% %5.1f %6d  L%-4s %s" % (seg, a & 0xfffff, 100.0 * inst / tot, thr / max(1, inst), smp, line, op))
    m = op.split()
    name = m[1] if m and m[0].startswith("@") and len(m) > 1 else (m[0] if m else "")
    if name.startswith("BAR"):
        seg += 1
