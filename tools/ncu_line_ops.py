#!/usr/bin/env python3
"""Per-source-line opcode histogram of one kernel from ncu's correlated source page.

  ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > cs.csv
  tools/ncu_line_ops.py cs.csv [top]

Prints the executed-instruction share of each source line together with the opcodes it compiled to,
then the opcode mix of the whole kernel (which pipe the work lands on: LOP3/SHF/ISETP/SEL/IADD3 = alu,
IMAD.* = fma, BSSY/BSYNC/BRA = branch unit)."""
import collections
import csv
import re
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
    hdr = rows[h]
    i_line, i_addr = hdr.index("Line No"), hdr.index("Address")
    i_src = [i for i, x in enumerate(hdr) if x == "Source"]
    i_exec = hdr.index("Instructions Executed")
    cur, src, fname = None, {}, ""
    per = collections.defaultdict(collections.Counter)
    mix = collections.Counter()
    tot = 0
    for r in rows[h + 1:]:
        if r and r[0] == "File Path":
            fname = r[1].rsplit("/", 1)[-1]
            continue
        if len(r) <= i_exec or r[i_line] == "Line No":
            continue
        if r[i_line]:
            cur = (fname, int(r[i_line]))
            src[cur] = r[i_src[0]].strip()
            continue
        if not r[i_addr].startswith("0x"):
            continue
        try:
            n = int(r[i_exec])
        except ValueError:
            continue
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[i_src[1]])
        full = m.group(2) if m else "?"
        op = full.split(".")[0]
        per[cur][op] += n
        mix[".".join(full.split(".")[:2]) if op == "IMAD" else op] += n
        tot += n
    for ln, c in sorted(per.items(), key=lambda kv: -sum(kv[1].values()))[:top]:
        t = sum(c.values())
        print(f"{ln[0][:22]:22s}:{ln[1]:<5d} {100 * t / tot:5.2f}%  {src.get(ln, '')[:100]}")
        print("             " + " ".join(f"{k}:{100 * v / tot:.2f}" for k, v in c.most_common(8)))
    print(f"\nwarp instructions executed: {tot}")
    for k, v in mix.most_common(30):
        print(f"  {k:12s} {100 * v / tot:6.2f}%")


if __name__ == "__main__":
    main()
