"""ctypes access to tests/emu/libmzemu.so: the product's kernel sources run on the CPU emulator."""
import ctypes as C
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class EmuLib:
    def __init__(self):
        self.lib = L = C.CDLL(os.path.join(ROOT, "tests/emu/libmzemu.so"))
        vp, u64, u32 = C.c_void_p, C.c_uint64, C.c_uint32
        L.emu_deflate.restype = C.c_int64
        L.emu_deflate.argtypes = [vp, u64, u32, C.c_int, u32, vp, u64, u32, vp]
        L.emu_crc32.restype = u32
        L.emu_crc32.argtypes = [vp, u64, u64, u32, vp]
        L.emu_crc32_combine.restype = u32
        L.emu_crc32_combine.argtypes = [u32, u32, u64]
        L.emu_inflate.restype = C.c_int32
        L.emu_inflate.argtypes = [vp, u64, vp, u64, u64, u64, C.POINTER(u64), C.POINTER(u64), C.POINTER(u32)]
        L.emu_inflate_spec.restype = C.c_int32
        L.emu_inflate_spec.argtypes = [vp, u64, vp, u64, u64, u32, u64, C.POINTER(u64), C.POINTER(u64), C.POINTER(u32)]
        L.emu_sha256.restype = None
        L.emu_sha256.argtypes = [vp, vp, vp, u32, vp]

    def deflate(self, data, level=1, chunk=65536, final=True, grid=0):
        data = bytes(data)
        cap = len(data) + (len(data) // 16384 + 2) * 64 + 1024
        out = C.create_string_buffer(cap)
        nch = max(1, (len(data) + chunk - 1) // chunk)
        lens = (C.c_uint32 * nch)()
        flags = final if (isinstance(final, int) and not isinstance(final, bool)) else (1 if final else 0)  # 1 = FINAL, 2 = DICT (one stream: chunks may refer back)
        n = self.lib.emu_deflate(data, len(data), chunk, level, flags, out, cap, grid, lens)
        if n < 0:
            raise ValueError(n)
        return out.raw[:n], list(lens)

    def crc32(self, data, seg=65536, misalign=0):
        data = bytes(data)
        nseg = (len(data) + seg - 1) // seg
        segs = (C.c_uint32 * max(nseg, 1))()
        v = self.lib.emu_crc32(data, len(data), seg, misalign, segs)
        return v, list(segs)[:nseg]

    def inflate(self, comp, out_cap, in_window=0, out_window=0):
        comp = bytes(comp)
        out = C.create_string_buffer(max(out_cap, 1) + 300)
        cons, prod, blocks = C.c_uint64(0), C.c_uint64(0), C.c_uint32(0)
        st = self.lib.emu_inflate(comp, len(comp), out, out_cap, in_window, out_window, C.byref(cons), C.byref(prod), C.byref(blocks))
        return st, out.raw[:prod.value], cons.value, blocks.value

    def inflate_spec(self, comp, out_cap, seg_bytes=2048, max_seg=64, window=0):
        """K6 rounds + serial fallback; returns (status, bytes, consumed, stats dict)"""
        comp = bytes(comp)
        out = C.create_string_buffer(max(out_cap, 1) + 300)
        cons, prod = C.c_uint64(0), C.c_uint64(0)
        stats = (C.c_uint32 * 5)()
        st = self.lib.emu_inflate_spec(comp, len(comp), out, out_cap, seg_bytes, max_seg, window, C.byref(cons), C.byref(prod), stats)
        names = ["rounds", "chain", "serial", "candidates", "discarded"]
        return st, out.raw[:prod.value], cons.value, dict(zip(names, list(stats)))

    def sha256(self, messages, align=16, shift=0):
        """digests of a list of byte strings packed at `align`-aligned offsets (+ shift) of one buffer"""
        offs, pos = [], shift
        for m in messages:
            offs.append(pos)
            pos += (len(m) + align - 1) // align * align + (0 if align > 1 else 0)
        buf = bytearray(pos + 64)
        for o, m in zip(offs, messages):
            buf[o:o + len(m)] = m
        n = len(messages)
        base = (C.c_uint8 * (len(buf) + 16))()
        a = (16 - C.addressof(base) % 16) % 16  # 16-byte aligned base so that aligned offsets take the fast path
        C.memmove(C.addressof(base) + a, bytes(buf), len(buf))
        off = (C.c_uint64 * max(n, 1))(*offs)
        ln = (C.c_uint64 * max(n, 1))(*[len(m) for m in messages])
        dig = (C.c_uint8 * (32 * max(n, 1)))()
        self.lib.emu_sha256(C.addressof(base) + a, off, ln, n, dig)
        raw = bytes(dig)
        return [raw[32 * i:32 * i + 32] for i in range(n)]
