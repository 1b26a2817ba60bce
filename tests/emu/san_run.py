"""Runs the product's kernel sources (CPU emulator build with AddressSanitizer + UBSan) over a few inputs.
Started by tests/test_emu_kernels.py::test_emu_kernels_under_sanitizers with libasan preloaded; exits non-zero on any
sanitizer report (halt_on_error) or wrong result. TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import sys
import zlib

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import datagen
import emushim

_orig = C.CDLL


def _patched(path, *a, **k):
    if str(path).endswith("libmzemu.so"):
        path = os.path.join(HERE, "libmzemu_san.so")
    return _orig(path, *a, **k)


C.CDLL = _patched
emu = emushim.EmuLib()
C.CDLL = _orig

data = datagen.mixed(260_000, 7) + datagen.random_bytes(3_000, 1)
for level in (1, 6):
    comp, _ = emu.deflate(data, level=level)
    assert zlib.decompress(comp, -15) == data
for d2, flags in ((data, 3), (bytes(70_000), 1), ((b"\x01\x00\x00" * 30_000)[:80_001], 3), (datagen.text_like(150_000, 8), 3)):
    comp, _ = emu.deflate(d2, level=6, final=flags)  # the history variant: one stream (DICT), runs and short periods (near sources in M0)
    assert zlib.decompress(comp, -15) == d2
comp, _ = emu.deflate(b"", level=1)
assert zlib.decompress(comp, -15) == b""
for n in (0, 1, 5, 4097, 70_000):
    d = datagen.random_bytes(n, n)
    v, _ = emu.crc32(d, 4096, 3)
    assert v == zlib.crc32(d)
co = zlib.compressobj(6, zlib.DEFLATED, -15)
comp = co.compress(data) + co.flush()
st, out, cons, _ = emu.inflate(comp, len(data), 3000, 66000)
assert st == 1 and out == data and cons == len(comp)
st, out, cons, stats = emu.inflate_spec(comp, len(data), seg_bytes=2048, max_seg=128)
assert st == 1 and out == data and stats["chain"] >= 3, stats
st, _, _, _ = emu.inflate_spec(comp[:len(comp) // 2], len(data), seg_bytes=2048, max_seg=128)
assert st == -5
bad = bytearray(comp)
for k in range(len(bad) // 3, len(bad) // 3 + 50):
    bad[k] ^= 0x3C
st, _, _, _ = emu.inflate_spec(bytes(bad), len(data) + 70_000, seg_bytes=2048, max_seg=128)
assert st != 0
print("sanitized run ok")
