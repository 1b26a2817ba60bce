#!/usr/bin/env python3
"""Group tools/ncu_by_line.py output by the phases of deflate_kernel.cuh (anchors = the phase comments)."""
import re, sys
byline, srcpath = sys.argv[1], sys.argv[2]
src = open(srcpath).read().splitlines()
def find(s): return next(i + 1 for i, l in enumerate(src) if s in l)
anchors = [('symbol map', 1), ('token_code', find('One token-list entry = one code word')), ('stage_put', find('OR `n` (<=32) bits')), ('BitWriter', find('struct BitWriter')),
           ('warp code build (unused)', find('warp-parallel code construction')), ('block_excl_sum', find('uint32_t block_excl_sum')), ('block_excl_max', find('uint32_t block_excl_max')),
           ('D block_build_codes', find('D: block-parallel literal')), ('header', find('Dynamic block header with a FIXED')), ('A extend/find_match', find('A: match finding')),
           ('kernel prologue/load', find('template <int WAYS, bool LAZY>')), ('A parse loop', find('---- A: parse')), ('B cover + T classify', find('---- B: cover')),
           ('T list write', find('---- T: write the ordered')), ('C matches', find('---- C (matches)')), ('C literals', find('---- C (literals)')), ('D call/wait', find('---- D: codes (10 warps')),
           ('E count', find('---- E: bit counts')), ('F emit', find('---- F: emit')), ('stored', find('stored block: header, pad')), ('G flush + trailer', find('---- G: flush')), ('end', len(src) + 1)]
rows = []
for l in open(byline):
    m = re.match(r'(\S+)\s*:\s*(\d+) inst\s+([\d.]+)% smp\s+([\d.]+)%', l)
    if m: rows.append((m.group(1), int(m.group(2)), float(m.group(3)), float(m.group(4))))
    elif l.startswith('total'): print(l.strip())
for (n, a), (_, b) in zip(anchors, anchors[1:]):
    sel = [r for r in rows if r[0].startswith('deflate_kernel') and a <= r[1] < b]
    print('%-26s lines %4d-%4d  instructions %6.2f%%  stall samples %6.2f%%' % (n, a, b - 1, sum(r[2] for r in sel), sum(r[3] for r in sel)))
for f, n in (('mzcuda_common', 'common (Smem accessors, scans)'), ('sm_3', 'intrinsics headers'), ('device_atomic', 'atomics header')):
    sel = [r for r in rows if r[0].startswith(f)]
    print('%-26s                  instructions %6.2f%%  stall samples %6.2f%%' % (n, sum(r[2] for r in sel), sum(r[3] for r in sel)))
