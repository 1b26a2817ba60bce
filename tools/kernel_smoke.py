#!/usr/bin/env python3
"""Every kernel once on small inputs, results checked against zlib: the workload for `compute-sanitizer --tool memcheck`
(tools/gpu/sanitize.sh). Not a benchmark."""
import ctypes as C
import os
import sys
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch

import __graft_entry__ as ge
import cuharness
import datagen

pkg = ge._load_pkg()
lib = pkg.load()
pkg.check(lib.mz_cuda_init())
torch.cuda.set_device(0)
tl = cuharness.TestLib()

# K2+K3, K1, K4 through the stream (all level profiles), reference-checked
data = datagen.mixed(3_000_000, 11) + datagen.random_bytes(70_001, 2) + datagen.text_like(1_000_000, 12)
for level in (0, 1, 2, 6, 9):
    comp, info = tl.compress(lib.mz_stream_cuda_create, data, level=level, window_bits=31, write_size=100_000)
    assert info["close"] == 0 and zlib.decompress(comp, 31) == data, level
    print("deflate level", level, "ok", len(comp), flush=True)
# K5 batch + K6 (medium member on the small windows, long member on the big ones) + serial path
for n, lvl in ((5_000_000, 6), (40_000_000, 6), (300_000, 1)):
    plain = datagen.text_like(n, seed=n % 97)
    co = zlib.compressobj(lvl, zlib.DEFLATED, 31)
    gz = co.compress(plain) + co.flush()
    for spec in ("1", "0") if n <= 5_000_000 else ("1",):
        os.environ["MZ_CUDA_SPEC"] = spec
        out, info = tl.decompress(lib.mz_stream_cuda_create, gz, n, window_bits=31, read_size=70_000)
        assert info["read"] == n and out == plain and info["total_in"] == len(gz), (n, spec, info)
    print("inflate", n, "ok", flush=True)
os.environ["MZ_CUDA_SPEC"] = "1"
# truncated and corrupted members must fail cleanly
out, info = tl.decompress(lib.mz_stream_cuda_create, gz[:len(gz) // 2], n, window_bits=31, read_size=70_000)
assert info["error"] != 0
big = datagen.text_like(12_000_000, seed=5)
co = zlib.compressobj(6, zlib.DEFLATED, 31)
gz = bytearray(co.compress(big) + co.flush())
for k in range(len(gz) // 2, len(gz) // 2 + 40):
    gz[k] ^= 0x77
out, info = tl.decompress(lib.mz_stream_cuda_create, bytes(gz), len(big), window_bits=31, read_size=1 << 20)
assert info["error"] != 0
# CRC through the replaced symbol (GPU path above 1 MiB)
buf = C.create_string_buffer(data, len(data))
assert lib.mz_crypt_crc32_update(0, buf, len(data)) == zlib.crc32(data)
print("kernel smoke ok", flush=True)
