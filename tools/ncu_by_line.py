#!/usr/bin/env python3
"""Join an ncu SASS source page with nvdisasm line info: per CUDA source line instruction counts / stall samples.

usage: ncu_by_line.py <report.ncu-rep> <cubin> <mangled kernel name> [top N]
"""
import csv, re, subprocess, sys
rep, cubin, kname = sys.argv[1:4]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 50
dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], stdout=subprocess.PIPE, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(".text." + kname + ":"))
off2line = {}
cur = ("?", 0)
for l in dis[start + 1:]:
    if l.startswith("//---") and ".text." in l:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        off2line[int(m.group(1), 16)] = (cur, m.group(2).strip())
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ia, ii, it, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
base = int(body[0][ia], 16)
agg = {}
tot_i = tot_s = 0
for r in body:
    off = int(r[ia], 16) - base
    (f, ln), _ = off2line.get(off, (("?", 0), ""))
    a = agg.setdefault((f, ln), [0, 0, 0, {}])
    a[0] += int(r[ii]); a[1] += int(r[it]); a[2] += int(r[isamp])
    for c in stall_cols:
        v = int(r[c] or 0)
        if v:
            a[3][hdr[c]] = a[3].get(hdr[c], 0) + v
    tot_i += int(r[ii]); tot_s += int(r[isamp])
print("total warp-instructions %d, samples %d" % (tot_i, tot_s))
src_cache = {}
def src(f, ln):
    import glob
    if f not in src_cache:
        c = glob.glob("/root/repo/minizip-ng_b200/csrc/" + f)
        src_cache[f] = open(c[0]).read().splitlines() if c else []
    L = src_cache[f]
    return L[ln - 1].strip()[:70] if 0 < ln <= len(L) else ""
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:topn]:
    st = sorted(a[3].items(), key=lambda kv: -kv[1])[:3]
    print("%-20s:%4d inst %5.2f%% smp %5.2f%% thr/inst %4.1f  %-70s %s" % (f, ln, 100 * a[0] / tot_i, 100 * a[2] / max(1, tot_s), a[1] / max(1, a[0]), src(f, ln),
          " ".join("%s=%d" % (k.replace("stall_", ""), v) for k, v in st)))
