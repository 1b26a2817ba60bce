#!/usr/bin/env python3
"""Compression ratio of the deflate kernel run on the CPU emulator (tests/emu) -- for trying parse / header changes
without a GPU round trip. usage: emu_ratio.py [MiB per dataset] [levels]"""
import os, sys, time, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import datagen, emushim, textgen
mib = float(sys.argv[1]) if len(sys.argv) > 1 else 2
levels = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 2, 6]
n = int(mib * (1 << 20))
E = emushim.EmuLib()
sets = {"text": textgen.host(n, seed=1), "records": datagen.binary_records(n, seed=5), "mix": datagen.mixed(8 << 20, seed=3)[:n],
        "zeros": bytes(n)}
for name, data in sets.items():
    for lv in levels:
        t = time.time()
        c, _ = E.deflate(data, level=lv)
        assert zlib.decompress(c, -15) == data
        z = zlib.compressobj(lv, zlib.DEFLATED, -15)
        zn = len(z.compress(data)) + len(z.flush())
        print("%-8s L%d ours %.4f zlib %.4f  (%.1fs)" % (name, lv, len(c) / len(data), zn / len(data), time.time() - t))
