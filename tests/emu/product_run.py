"""Scenarios for tests/test_emu_product.py: the WHOLE product library (real host C + real API shim + kernel sources) built
against the CPU execution-model emulator (tests/emu/libmz_strm_emu.so) and driven through the vtbl exactly like the GPU
tests drive libmz_strm_cuda.so. One scenario per process (workspaces are pooled per process and read their environment
knobs when created). TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import sys
import zlib

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import cuharness
import datagen
import refshim

p = cuharness.pkg()
lib = p.configure(C.CDLL(os.environ.get("MZ_EMU_LIB") or os.path.join(HERE, "libmz_strm_emu.so")))  # MZ_EMU_LIB: the sanitizer build
tl = cuharness.TestLib()
CREATE = lib.mz_stream_cuda_create


def gz(data, level=6, wbits=31, flush_every=0):
    co = zlib.compressobj(level, zlib.DEFLATED, wbits)
    if not flush_every:
        return co.compress(data) + co.flush()
    parts = []
    for o in range(0, len(data), flush_every):
        parts.append(co.compress(data[o:o + flush_every]))
        parts.append(co.flush(zlib.Z_FULL_FLUSH))
    parts.append(co.flush())
    return b"".join(parts)


def scenario_write():
    """write path: multi-batch pipeline (tiny batches), every framing, odd write sizes; zlib and the reference read it back"""
    ref = refshim.RefLib() if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle", "_ref", "libmzref.so")) else None
    data = datagen.mixed(700_000, 5) + datagen.random_bytes(70_001, 1) + datagen.text_like(200_000, 2)
    for level, wbits, wsize in ((1, -15, 16384), (6, 31, 65535), (9, 15, 1), (0, 31, 100_000), (-1, -15, 333_333)):
        d = data if wsize > 1 else data[:5000]
        comp, info = tl.compress(CREATE, d, level=level, window_bits=wbits, write_size=wsize)
        assert info["close"] == 0 and info["total_in"] == len(d) and info["total_out"] == len(comp) == info["sink_tell"], info
        assert zlib.decompress(comp, wbits) == d, (level, wbits)
        if wbits == 31:
            assert comp[:4] == b"\x1f\x8b\x08\x00" and comp[8] == (2 if level == 9 else 4 if level in (0, 1) else 0)
            assert int.from_bytes(comp[-4:], "little") == len(d) and int.from_bytes(comp[-8:-4], "little") == zlib.crc32(d)
        if ref is not None:
            out, rinfo = ref.decompress_with(ref.lib.mz_stream_zlib_create, comp, window_bits=wbits, read_size=65536)
            assert rinfo["read_err"] == 0 and out == d and rinfo["total_in"] == len(comp)
    comp, info = tl.compress(CREATE, b"", level=6, window_bits=-15)
    assert comp == b"\x03\x00" and info["total_out"] == 2  # byte-identical to zlib
    comp, info = tl.compress(CREATE, b"", level=6, window_bits=31)
    assert comp == bytes.fromhex("1f8b0800000000000003") + b"\x03\x00" + bytes(8)


def scenario_read(spec):
    """read path on the small windows: foreign members, tiny reads, trailing garbage, TOTAL_IN_MAX, errors"""
    text = datagen.text_like(900_000, 7)
    mixed = datagen.mixed(500_000, 8) + datagen.random_bytes(90_000, 3) + text[:200_000]
    for plain, level, wbits, rsize, flush in ((text, 6, 31, 16384, 0), (mixed, 1, -15, 65535, 0), (mixed, 9, 15, 4096, 0), (text, 6, -15, 300_000, 30_000),
                                              (text[:70_000], 0, 31, 1000, 0), (b"", 6, 31, 10, 0), (b"a", 6, -15, 1, 0)):
        comp = gz(plain, level, wbits, flush)
        out, info = tl.decompress(CREATE, comp + b"\xee" * 5000 if wbits == -15 else comp, len(plain), window_bits=wbits, read_size=rsize)
        assert info["read"] == len(plain) and out == plain and info["read_again"] == 0, (level, wbits, info)
        assert info["total_in"] == len(comp) and info["total_out"] == len(plain) and info["error"] == 0 and info["close"] == 0, info
    comp = gz(text, 6, -15)
    out, info = tl.decompress(CREATE, comp + b"\x55" * 999, len(text), window_bits=-15, read_size=16384, total_in_max=len(comp))
    assert out == text and info["total_in"] == len(comp) and info["base_tell"] == len(comp)  # never reads past TOTAL_IN_MAX
    # truncated: clean prefix, then MZ_BUF_ERROR, sticky, close -> MZ_CLOSE_ERROR
    out, info = tl.decompress(CREATE, comp[:len(comp) // 2], len(text), window_bits=-15, read_size=16384)
    assert info["error"] == p.MZ_BUF_ERROR and info["close"] == p.MZ_CLOSE_ERROR, info
    # wrong gzip trailer
    g = bytearray(gz(text, 6, 31))
    g[-6] ^= 1
    out, info = tl.decompress(CREATE, bytes(g), len(text), window_bits=31, read_size=16384)
    assert info["error"] == p.MZ_DATA_ERROR and info["close"] == p.MZ_CLOSE_ERROR, info
    # raw deflate fed to a gzip reader
    out, info = tl.decompress(CREATE, comp, len(text), window_bits=31, read_size=16384)
    assert info["error"] == p.MZ_DATA_ERROR
    # corruption in the middle
    bad = bytearray(gz(text, 6, 31))
    for k in range(len(bad) // 2, len(bad) // 2 + 30):
        bad[k] ^= 0x5A
    out, info = tl.decompress(CREATE, bytes(bad), len(text), window_bits=31, read_size=16384)
    assert info["error"] != 0


def scenario_long():
    """a member that overflows the first window: the workspace switches to the long-stream windows and runs K6 rounds with the
    delivery / next-round overlap; output windows slide several times"""
    text = datagen.text_like(2_400_000, 11) + datagen.random_bytes(200_000, 4) + datagen.text_like(900_000, 12)
    for level, rsize in ((6, 65536), (1, 1 << 20)):
        comp = gz(text, level, 31)
        assert len(comp) > (1 << 20)
        out, info = tl.decompress(CREATE, comp, len(text), window_bits=31, read_size=rsize)
        assert info["read"] == len(text) and out == text and info["total_in"] == len(comp) and info["close"] == 0, info
    # abandoned half way, then the pooled workspace serves another stream
    src, keep = tl.source(comp)
    s = CREATE()
    tl.lib.mzt_set_prop(s, p.MZ_STREAM_PROP_COMPRESS_WINDOW, 31)
    tl.lib.mzt_set_base(s, src)
    assert tl.lib.mzt_open(s, None, p.MZ_OPEN_MODE_READ) == 0
    buf = C.create_string_buffer(50_000)
    assert tl.lib.mzt_read(s, buf, 50_000) > 0
    assert tl.lib.mzt_close(s) == 0
    tl.delete(s)
    tl.delete(src)
    out, info = tl.decompress(CREATE, comp, len(text), window_bits=31, read_size=100_000)
    assert out == text and info["total_in"] == len(comp)


def scenario_ring():
    """the compressed window as a ring with read-ahead (MZ_CUDA_READ_AHEAD=2: bytes are pulled from base behind the window whether
    or not a round is in flight -- on the emulator a launch has always finished): many wraps of a small ring, the trailer found
    across a wrap, TOTAL_IN_MAX honoured by the read-ahead (bytes behind the member are never touched), truncation reported"""
    small = os.environ.get("MZ_TEST_RING_SMALL") == "1"  # (the sanitizer run: no wrap, but the read-ahead accounting under ASan)
    text = (datagen.text_like(2_000_000, 21) + datagen.random_bytes(200_000, 5) + datagen.text_like(1_000_000, 22)) if small else \
           (datagen.text_like(4_500_000, 21) + datagen.random_bytes(300_000, 5) + datagen.text_like(2_000_000, 22))
    # 262146 = 4 * 65536 + 2: a copy the helper threads split into four 64 KiB slices -- plus two bytes that belong to the last one
    for level, rsize in ((6, 262146), (1, 65536)):
        comp = gz(text, level, 31)
        assert len(comp) > ((1 << 20) if small else (5 << 19))
        out, info = tl.decompress(CREATE, comp, len(text), window_bits=31, read_size=rsize)
        assert info["read"] == len(text) and out == text and info["total_in"] == len(comp) and info["close"] == 0, info
        if level == 1:
            continue  # (the variations below once, on the level-6 member)
        junk = comp + datagen.random_bytes(1_500_000, 6)
        out, info = tl.decompress(CREATE, junk, len(text), window_bits=31, read_size=rsize, total_in_max=len(comp))
        assert out == text and info["total_in"] == len(comp) and info["base_tell"] == len(comp) and info["close"] == 0, info
        out, info = tl.decompress(CREATE, comp[:len(comp) - 100_000], len(text), window_bits=31, read_size=rsize)
        assert info["error"] == p.MZ_BUF_ERROR and info["read_again"] == p.MZ_BUF_ERROR and info["total_out"] > len(text) // 2, info


def scenario_crc():
    data = datagen.random_bytes(3_000_001, 9)
    buf = C.create_string_buffer(data, len(data))
    assert lib.mz_crypt_crc32_update(0, buf, len(data)) == zlib.crc32(data)  # above the host threshold: K1 + fold
    half = len(data) // 2
    v = lib.mz_crypt_crc32_update(0, buf, half)
    v = lib.mz_crypt_crc32_update(v, C.byref(buf, half), len(data) - half)
    assert v == zlib.crc32(data)
    # the host path (slice-by-16, calls below the threshold): every length 0..70, odd alignments, chaining, known answers
    assert lib.mz_crypt_crc32_update(0, C.create_string_buffer(b"123456789", 9), 9) == 0xCBF43926
    assert lib.mz_crypt_crc32_update(0x1234, buf, 0) == 0x1234  # size 0 returns value unchanged (mz_os.c:340)
    for n in list(range(0, 71)) + [255, 256, 257, 4095, 65535]:
        for a in (0, 1, 3, 7, 13):
            want = zlib.crc32(data[a:a + n])
            assert lib.mz_crypt_crc32_update(0, C.byref(buf, a), n) == want, (n, a)
            k = n // 3
            v = lib.mz_crypt_crc32_update(0, C.byref(buf, a), k)
            assert lib.mz_crypt_crc32_update(v, C.byref(buf, a + k), n - k) == want, (n, a, "chained")
    # pieces: more than one 8 MiB staging piece, uneven tail
    big = datagen.random_bytes(19_000_003, 10)
    bb = C.create_string_buffer(big, len(big))
    assert lib.mz_crypt_crc32_update(0, bb, len(big)) == zlib.crc32(big)
    assert lib.mz_crypt_crc32_update(0xDEADBEEF, bb, len(big)) == zlib.crc32(big, 0xDEADBEEF)


def scenario_sharded():
    """mz_cuda_deflate_sharded with three shards (all on the emulator's one device): region layout, piece pipeline, rows, CRC fold"""
    pkg = p
    text = datagen.text_like(11 * 65536 - 4321, 77)
    cuts = [0, 5 * 65536, 8 * 65536, len(text)]
    n = 3
    bound = [lib.mz_cuda_gather_region_bound(cuts[i + 1] - cuts[i]) for i in range(n)]
    cap = sum(bound)
    nch_all = sum((cuts[i + 1] - cuts[i] + 65535) // 65536 for i in range(n))
    ins = [C.create_string_buffer(text[cuts[i]:cuts[i + 1]] + bytes(64), cuts[i + 1] - cuts[i] + 64) for i in range(n)]
    gath = [C.create_string_buffer(cap + 64) for _ in range(n)]
    rows = [(C.c_uint32 * (3 * nch_all))() for _ in range(n)]

    def al(b):  # 16-byte aligned view into a ctypes buffer
        a = C.addressof(b)
        return (a + 15) & ~15
    # ctypes buffers are not 16-byte aligned by contract: copy into aligned storage
    store = []

    def aligned(data, extra=64):
        raw = C.create_string_buffer(len(data) + extra + 16)
        p = al(raw)
        C.memmove(p, data, len(data))
        store.append(raw)
        return p
    shards = (pkg.Shard * n)()
    gp = []
    for i in range(n):
        shards[i].device = 0
        shards[i].d_in = aligned(text[cuts[i]:cuts[i + 1]])
        shards[i].len = cuts[i + 1] - cuts[i]
        g = aligned(bytes(cap), 0)
        gp.append(g)
        shards[i].d_gathered = g
        shards[i].gathered_cap = cap
        shards[i].d_rows = C.addressof(rows[i])
    roff = (C.c_uint64 * n)()
    slen = (C.c_uint64 * n)()
    crc = C.c_uint32(0)
    for pieces in (1, 2, 3):
        err = lib.mz_cuda_deflate_sharded(shards, n, 1, pieces, roff, slen, C.byref(crc))
        assert err == 0, (err, lib.mz_cuda_last_error())
        assert crc.value == zlib.crc32(text)
        streams = []
        for j in range(n):  # every device's gathered buffer holds every stream at the same place
            parts = [C.string_at(gp[j] + roff[i], slen[i]) for i in range(n)]
            streams.append(b"".join(parts))
        assert streams[0] == streams[1] == streams[2]
        d = zlib.decompressobj(-15)
        assert d.decompress(streams[0]) == text and d.eof and d.unused_data == b""
        # rows: {crc32, in_len, out_len} per chunk in global order, identical everywhere
        r0 = list(rows[0])
        assert list(rows[1]) == r0 and list(rows[2]) == r0
        pos = 0
        total_out = 0
        for c in range(nch_all):
            ln = r0[3 * c + 1]
            assert r0[3 * c] == zlib.crc32(text[pos:pos + ln])
            pos += ln
            total_out += r0[3 * c + 2]
        assert pos == len(text) and total_out == sum(slen)


if __name__ == "__main__":
    name = sys.argv[1]
    if name == "write":
        scenario_write()
    elif name == "read":
        scenario_read(os.environ.get("MZ_CUDA_SPEC", "1"))
    elif name == "long":
        scenario_long()
    elif name == "ring":
        scenario_ring()
    elif name == "crc":
        scenario_crc()
    elif name == "sharded":
        scenario_sharded()
    else:
        raise SystemExit("unknown scenario " + name)
    print("scenario %s ok" % name)
