#!/usr/bin/env python3
"""Split an ncu SASS source page of one kernel into segments at barriers / selected marker opcodes and report per segment
warp-instructions, lane utilisation, stall samples, shared-memory wavefronts and the share of integer-ALU-pipe opcodes.

usage: ncu_segments.py <report.ncu-rep> [kernel-substring]
Segments are in address order; each line shows the first source opcode after the barrier so it can be matched to the
kernel's phase structure.  Inlined helpers are attributed to the segment that executes them, which a
per-source-line view (tools/ncu_by_line.py) cannot do."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
ALU = ("LOP3", "SHF", "SEL", "ISETP", "IADD3", "VIADD", "PRMT", "LEA", "VIMNMX", "VIADDMNMX", "IABS", "MOV", "P2R", "R2P", "PLOP3", "BMSK", "SGXT", "FLO", "BREV", "POPC", "CS2R", "IMNMX", "LOP")
XU = ("FLO", "BREV", "POPC", "MUFU")
segs = []
cur = {"start": 0, "inst": 0, "thr": 0, "smp": 0, "wf": 0, "wfi": 0, "alu": 0, "lsu": 0, "first": ""}
tot = 0
for k, r in enumerate(body):
    op = r[col["Source"]].strip()
    m = op.split()
    name = m[1] if m and m[0].startswith("@") and len(m) > 1 else (m[0] if m else "")
    inst = int(r[col["Instructions Executed"]])
    cur["inst"] += inst
    cur["thr"] += int(r[col["Thread Instructions Executed"]])
    cur["smp"] += int(r[col["# Samples"]])
    cur["wf"] += int(r[col["L1 Wavefronts Shared"]] or 0)
    cur["wfi"] += int(r[col["L1 Wavefronts Shared Ideal"]] or 0)
    if name.startswith(ALU) and not name.startswith("IMAD"):
        cur["alu"] += inst
    if name.startswith(("LDS", "STS", "ATOMS", "LDG", "STG", "SHFL", "RED", "UBLKCP", "LDL", "STL")):
        cur["lsu"] += inst
    tot += inst
    if name.startswith("BAR") or name.startswith("CALL") and False:
        cur["end"] = k
        segs.append(cur)
        cur = {"start": k + 1, "inst": 0, "thr": 0, "smp": 0, "wf": 0, "wfi": 0, "alu": 0, "lsu": 0, "first": ""}
cur["end"] = len(body) - 1
segs.append(cur)
tsmp = sum(s["smp"] for s in segs)
twf = sum(s["wf"] for s in segs)
print("total warp-instructions %d  samples %d  shared wavefronts %d" % (tot, tsmp, twf))
print("%4s %7s %7s %7s %6s %6s %6s %7s %7s  %s" % ("seg", "sass#", "inst%", "smp%", "lanes", "alu%", "lsu%", "wf%", "wf/ideal", "first opcodes"))
for i, s in enumerate(segs):
    if s["inst"] == 0:
        continue
    first = "; ".join(body[j][col["Source"]].strip()[:28] for j in range(s["start"], min(s["start"] + 2, len(body))))
    print("%4d %7d %6.2f%% %6.2f%% %6.1f %5.1f%% %5.1f%% %6.2f%% %7.2f  %s" % (i, s["end"] - s["start"] + 1, 100 * s["inst"] / tot, 100 * s["smp"] / max(1, tsmp), s["thr"] / max(1, s["inst"]),
          100 * s["alu"] / max(1, s["inst"]), 100 * s["lsu"] / max(1, s["inst"]), 100 * s["wf"] / max(1, twf), s["wf"] / max(1, s["wfi"]), first))
