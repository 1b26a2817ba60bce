#!/usr/bin/env python3
"""Compression-ratio table: ours (GPU, 64 KiB chunks) vs zlib at the same nominal level (the reference's codec; whole stream
and cut into the same 64 KiB chunks) on the bench text, binary records and the C4 mix; also checks that two GPU runs give
byte-identical output (the encoder is deterministic).   usage: ratio_table.py [MiB per dataset]  -> markdown on stdout"""
import os, sys, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import __graft_entry__ as ge
import datagen, textgen
pkg = ge._load_pkg(); lib = pkg.load(); pkg.check(lib.mz_cuda_init()); torch.cuda.set_device(0)
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = mib << 20
sets = {"bench text (tests/support/textgen)": textgen.host(n, seed=1),
        "binary records (48-byte records with counters)": datagen.binary_records(n, seed=5),
        "C4 mix (70 % text, 20 % records, 10 % random)": (datagen.mixed(8 << 20, seed=3) * (n // (8 << 20) + 1))[:n]}


def ours(data, level, one_stream=False):
    src = torch.from_numpy(np.frombuffer(data, dtype=np.uint8).copy()).cuda()
    b = pkg.DeflateBatch(len(data))
    k = b.compress(src, len(data), level=level, final=True, one_stream=one_stream)
    joined, crc = b.result(k)
    return bytes(joined.cpu().numpy().tobytes()), crc


def zl(data, level, chunk=None):
    if not chunk:
        co = zlib.compressobj(level, zlib.DEFLATED, -15)
        return len(co.compress(data)) + len(co.flush())
    t = 0
    for o in range(0, len(data), chunk):
        co = zlib.compressobj(level, zlib.DEFLATED, -15)
        t += len(co.compress(data[o:o + chunk])) + len(co.flush())
    return t


print("| data (%d MiB) | level | ours, independent 64 KiB chunks | ours, one stream (MZ_CUDA_FLAG_DICT) | zlib whole stream | zlib 64 KiB chunks | ours one stream / zlib whole | ours chunks / zlib chunks | identical on re-run |" % mib)
print("|---|---|---|---|---|---|---|---|---|")
for name, data in sets.items():
    for level in (1, 2, 6, 9):
        c1, crc = ours(data, level)
        c2, _ = ours(data, level)
        c3, _ = ours(data, level, one_stream=True)
        assert zlib.decompress(c1, -15) == data and (crc & 0xffffffff) == zlib.crc32(data) and zlib.decompress(c3, -15) == data
        zw, zc = zl(data, level), zl(data, level, 65536)
        print("| %s | %d | %.4f | %.4f | %.4f | %.4f | %.3f | %.3f | %s |" % (name, level, len(c1) / n, len(c3) / n, zw / n, zc / n, len(c3) / zw, len(c1) / zc, "yes" if c1 == c2 else "NO"))
