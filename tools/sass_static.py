#!/usr/bin/env python3
"""Static SASS census of one kernel of the built library, split at barriers: instructions per segment by pipe class
(ALU / FMA(IMAD) / LSU / other).  A no-GPU proxy for the dynamic instruction count of straight-line unrolled phases.
usage: sass_static.py [kernel-substring] [lib]"""
import os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pat = sys.argv[1] if len(sys.argv) > 1 else "deflate_chunks_kernelILi2ELb0"
lib = sys.argv[2] if len(sys.argv) > 2 else os.path.join(root, "minizip-ng_b200", "libmz_strm_cuda.so")
out = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True).stdout
ALU = ("LOP3", "SHF", "SEL", "ISETP", "IADD3", "VIADD", "PRMT", "LEA", "VIMNMX", "VIADDMNMX", "IABS", "MOV", "P2R", "R2P", "PLOP3", "BMSK", "SGXT", "CS2R", "IMNMX", "LOP", "FLO", "BREV", "POPC")
LSU = ("LDS", "STS", "ATOMS", "LDG", "STG", "SHFL", "RED", "UBLKCP", "LDL", "STL", "ATOMG")
infn = False
segs = [[0, 0, 0, 0]]
for line in out.splitlines():
    if "Function :" in line:
        infn = pat in line
        continue
    if not infn:
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if not m:
        continue
    toks = m.group(2).split()
    name = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    s = segs[-1]
    if name.startswith("IMAD") or name.startswith("FFMA") or name.startswith("FMUL"):
        s[1] += 1
    elif name.startswith(ALU):
        s[0] += 1
    elif name.startswith(LSU):
        s[2] += 1
    else:
        s[3] += 1
    if name.startswith("BAR"):
        segs.append([0, 0, 0, 0])
tot = [sum(s[i] for s in segs) for i in range(4)]
print("seg   alu   fma   lsu other total")
for i, s in enumerate(segs):
    if sum(s) > 20:
        print("%3d %5d %5d %5d %5d %5d" % (i, s[0], s[1], s[2], s[3], sum(s)))
print("all %5d %5d %5d %5d %5d" % (tot[0], tot[1], tot[2], tot[3], sum(tot)))
