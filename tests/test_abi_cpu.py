"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol the headers declare, and the
host logic that needs no GPU behaves like the reference (mz_strm_zlib.c:312-378). No compute calls here."""
import ctypes as C
import os
import re

import refshim


def _pkg(built):
    import cuharness
    return cuharness.pkg()


def _declared(header):
    txt = open(header).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mz_(?:stream_cuda|cuda|crypt|zip_cuda)_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(built):
    p = _pkg(built)
    lib = C.CDLL(p.LIB_PATH)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    names = (_declared(os.path.join(root, "include/mz_strm_cuda.h")) + _declared(os.path.join(root, "include/mz_cuda_batch.h")) +
             _declared(os.path.join(root, "include/mz_zip_cuda.h")))
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), "missing export " + n
    assert sorted(set(names)) == sorted(set(p.EXPORTS))


def test_vtbl_slot_order_and_object_header(built):
    p = _pkg(built)
    lib = p.load()
    vt = lib.mz_stream_cuda_get_interface()
    slots = (C.c_void_p * 12).from_address(vt)
    names = ["open", "is_open", "read", "write", "tell", "seek", "close", "error", "create", "delete", "get_prop_int64", "set_prop_int64"]
    for i, n in enumerate(names):  # mz_strm.h:53-67 order
        fn = C.cast(getattr(lib, "mz_stream_cuda_" + n), C.c_void_p).value
        assert slots[i] == fn, n
    s = lib.mz_stream_cuda_create()
    hdr = (C.c_void_p * 2).from_address(s)
    assert hdr[0] == vt and hdr[1] is None  # {vtbl, base} first (mz_strm.h:69-72)
    ps = C.c_void_p(s)
    lib.mz_stream_cuda_delete(C.byref(ps))
    assert ps.value is None


def test_props_and_lifecycle_without_gpu(built):
    p = _pkg(built)
    lib = p.load()
    s = lib.mz_stream_cuda_create()
    v = C.c_int64(0)
    assert lib.mz_stream_cuda_is_open(s) == refshim.MZ_OPEN_ERROR
    assert lib.mz_stream_cuda_set_prop_int64(s, refshim.PROP_COMPRESS_LEVEL, -1) == 0
    assert lib.mz_stream_cuda_set_prop_int64(s, refshim.PROP_COMPRESS_WINDOW, 31) == 0
    assert lib.mz_stream_cuda_set_prop_int64(s, refshim.PROP_TOTAL_IN_MAX, 1234) == 0
    assert lib.mz_stream_cuda_get_prop_int64(s, refshim.PROP_TOTAL_IN_MAX, C.byref(v)) == 0 and v.value == 1234
    assert lib.mz_stream_cuda_get_prop_int64(s, refshim.PROP_COMPRESS_WINDOW, C.byref(v)) == 0 and v.value == 31
    assert lib.mz_stream_cuda_get_prop_int64(s, refshim.PROP_TOTAL_OUT_MAX, C.byref(v)) == refshim.MZ_EXIST_ERROR
    assert lib.mz_stream_cuda_set_prop_int64(s, refshim.PROP_COMPRESS_METHOD, 8) == refshim.MZ_EXIST_ERROR
    assert lib.mz_stream_cuda_tell(s) == refshim.MZ_TELL_ERROR
    assert lib.mz_stream_cuda_seek(s, 0, 0) == refshim.MZ_SEEK_ERROR
    # invalid parameters are rejected before the GPU is even looked at (zip_fuzzer.c feeds arbitrary levels)
    lib.mz_stream_cuda_set_prop_int64(s, refshim.PROP_COMPRESS_LEVEL, 77)
    assert lib.mz_stream_cuda_open(s, None, refshim.MZ_OPEN_MODE_WRITE) == refshim.MZ_OPEN_ERROR
    assert lib.mz_stream_cuda_close(s) == refshim.MZ_CLOSE_ERROR  # latched error, like zlib->error (:302-304)
    ps = C.c_void_p(s)
    lib.mz_stream_cuda_delete(C.byref(ps))


def test_no_gpu_means_support_error_not_a_fallback(built):
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("GPU present")
    p = _pkg(built)
    lib = p.load()
    s = lib.mz_stream_cuda_create()
    assert lib.mz_stream_cuda_open(s, None, refshim.MZ_OPEN_MODE_WRITE) == refshim.MZ_SUPPORT_ERROR
    assert lib.mz_stream_cuda_is_open(s) == refshim.MZ_OPEN_ERROR
    ps = C.c_void_p(s)
    lib.mz_stream_cuda_delete(C.byref(ps))
    assert lib.mz_cuda_init() == refshim.MZ_SUPPORT_ERROR


def test_small_crc_calls_match_reference_semantics(built, orc):
    """Below the size threshold mz_crypt_crc32_update is the mz_crypt.c:81-90 loop: byte-at-a-time chaining (pkcrypt)."""
    p = _pkg(built)
    lib = p.load()
    data = b"The quick brown fox jumps over the lazy dog"
    v = 0
    for b in data:
        v = lib.mz_crypt_crc32_update(v, bytes([b]), 1)
    assert v == orc.crc32(0, data) == 0x414FA339
    assert lib.mz_crypt_crc32_update(5, None, 0) == 5


def test_host_crc_paths_and_clmul_constants(built, orc):
    """Calls below the GPU threshold: the carry-less-multiplication path (x86-64 with PCLMULQDQ; 64 bytes and more) and the table
    loop must both give zlib's value for every length, alignment and start value; the folding constants written into
    mz_crypt_cuda.c are re-derived here from the polynomial (k(n) = bitreflect32(x^n mod P) << 1, mu', P')."""
    import random
    import subprocess
    import sys
    P = 0x104C11DB7

    def xpow(n):
        r = 1
        for _ in range(n):
            r <<= 1
            if r >> 32:
                r ^= P
        return r

    def refl(v, bits):
        return int(bin(v)[2:].zfill(bits)[::-1], 2)

    num, q = 1 << 64, 0
    for i in range(64, 31, -1):
        if (num >> i) & 1:
            q |= 1 << (i - 32)
            num ^= P << (i - 32)
    want = {"k512": (refl(xpow(4 * 128 + 32), 32) << 1, refl(xpow(4 * 128 - 32), 32) << 1), "k128": (refl(xpow(128 + 32), 32) << 1, refl(xpow(128 - 32), 32) << 1),
            "k64": (refl(xpow(64), 32) << 1, 0), "pmu": (refl(P, 33), refl(q, 33))}
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "minizip-ng_b200", "csrc", "mz_crypt_cuda.c")).read()
    for name, (lo, hi) in want.items():
        m = re.search(r"const __m128i %s = _mm_set_epi64x\((0x[0-9a-f]+|0)(?:ll)?, (0x[0-9a-f]+)ll\)" % name, src)
        assert m and int(m.group(1), 16) == hi and int(m.group(2), 16) == lo, (name, hex(lo), hex(hi), m and m.groups())
    # both host paths against zlib (a child process each: the choice is latched at the first call)
    code = r"""
import ctypes as C, os, random, sys, zlib
lib = C.CDLL(sys.argv[1]); lib.mz_crypt_crc32_update.restype = C.c_uint32; lib.mz_crypt_crc32_update.argtypes = [C.c_uint32, C.c_void_p, C.c_int32]
rng = random.Random(int(sys.argv[2])); data = bytes(rng.randrange(256) for _ in range(200000)); buf = C.create_string_buffer(data, len(data) + 64)
for it in range(6000):
    a = rng.randrange(0, 64); n = rng.choice([rng.randrange(0, 400), rng.randrange(0, 150000), 63, 64, 65, 79, 80, 81, 127, 128, 129, 4096, 65535])
    n = min(n, len(data) - a); v0 = rng.choice([0, 0xffffffff, rng.randrange(1 << 32)])
    assert lib.mz_crypt_crc32_update(v0, C.byref(buf, a), n) == zlib.crc32(data[a:a + n], v0), (a, n, v0)
print("host crc ok")
"""
    p = _pkg(built)
    for env in ({}, {"MZ_CUDA_CRC_NO_CLMUL": "1"}):
        r = subprocess.run([sys.executable, "-c", code, p.LIB_PATH, "7"], env=dict(os.environ, **env), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=300)
        assert r.returncode == 0 and b"host crc ok" in r.stdout, r.stdout[-2000:]


def test_crc_combine_host_arithmetic(built, orc):
    p = _pkg(built)
    lib = p.load()
    import datagen
    a, b = datagen.random_bytes(1000, 1), datagen.random_bytes(123457, 2)
    assert lib.mz_cuda_crc32_combine(orc.crc32(0, a), orc.crc32(0, b), len(b)) == orc.crc32(0, a + b)
    for ln in (0, 1, 65536, (1 << 32) + 17):
        assert lib.mz_cuda_crc32_combine(0xDEADBEEF, 0x01020304, ln) == orc.crc32_combine(0xDEADBEEF, 0x01020304, ln)


def test_mem64_support_stream(built):
    import cuharness
    tl = cuharness.TestLib()
    sink = tl.sink()
    buf = C.create_string_buffer(b"abcdef", 6)
    assert tl.lib.mzt_write(sink, buf, 6) == 6 and tl.lib.mzt_tell(sink) == 6
    assert tl.sink_bytes(sink) == b"abcdef"
    tl.delete(sink)
    # the multi-threaded copy of large writes lands the same bytes as the plain memcpy
    import datagen
    blob = datagen.random_bytes((9 << 20) + 12345, seed=9)
    src = C.create_string_buffer(blob, len(blob))
    for threads in (1, 8):
        dst = C.create_string_buffer(len(blob) + 64)
        sink = tl.sink()
        tl.lib.mz_stream_mem64_set_sink(sink, dst, len(blob) + 64)
        tl.lib.mz_stream_mem64_set_copy_threads(sink, threads)
        assert tl.lib.mzt_write_all(sink, src, len(blob), len(blob)) == len(blob) and tl.lib.mzt_tell(sink) == len(blob)
        assert dst.raw[:len(blob)] == blob
        tl.delete(sink)


def test_zip_batch_calls_need_the_reference_container(built):
    """mz_zip_cuda_* take the container functions from the HOST program (weak references); in a process without
    mz_zip.c they must say so instead of crashing, before touching any GPU state. The file-info mirror has the
    layout of mz_zip_file on this ABI (mz_zip.h:25-51: 3 time_t, 2 int64, 4 pointers ...)."""
    p = _pkg(built)
    lib = C.CDLL(p.LIB_PATH)
    lib.mz_zip_cuda_add_buffers.restype = C.c_int32
    lib.mz_zip_cuda_add_buffers.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int16, C.c_void_p]
    lib.mz_zip_cuda_extract_all.restype = C.c_int32
    lib.mz_zip_cuda_extract_all.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    fake = C.create_string_buffer(64)
    assert lib.mz_zip_cuda_add_buffers(fake, None, 0, 6, None) == p.MZ_SUPPORT_ERROR
    assert lib.mz_zip_cuda_extract_all(fake, None, None, None) == p.MZ_SUPPORT_ERROR
    assert lib.mz_zip_cuda_abi_file_info_size() == 128  # == sizeof(mz_zip_file); the C4 driver asserts the same against the reference header
