"""The product's kernel SOURCES executed on the CPU emulator (tests/emu): logic parity without a GPU.

Same bar as the GPU tests (bit-exact vs the oracle / system zlib) at sizes the emulator finishes in seconds.
This is a debugging aid for the build container, not a product path: nothing here is shipped or measured.
"""
import os
import subprocess
import zlib

import pytest

import datagen

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emu():
    r = subprocess.run(["make", "-s"], cwd=os.path.join(HERE, "emu"), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    import emushim
    return emushim.EmuLib()


@pytest.mark.parametrize("level", [0, 1, 4, 6])
def test_emu_deflate_roundtrip(emu, orc, level):
    for kind, n in (("text", 70001), ("mixed", 65536), ("rand", 33000), ("zeros", 140000), ("text", 0), ("text", 1), ("abc", 32769)):
        data = {"text": datagen.text_like(n, 3), "mixed": datagen.mixed(n, 4), "rand": datagen.random_bytes(n, 5), "zeros": bytes(n),
                "abc": (b"abcabcabd" * (n // 9 + 1))[:n]}[kind]
        comp, lens = emu.deflate(data, level=level)
        assert zlib.decompress(comp, -15) == data
        err, out, cons = orc.inflate(comp, n + 8)
        assert err == 0 and out == data and cons == len(comp)


def test_emu_deflate_size_sweep(emu):
    """every size around the places where the kernel changes behaviour: empty and tiny chunks, the 258-byte match limit, the
    32 KiB sub-block boundary, the 64 KiB chunk limit, and ragged last chunks of a multi-chunk stream (sync markers between)"""
    base = datagen.text_like(200_000, 17)
    sizes = list(range(0, 41)) + list(range(254, 262)) + list(range(32760, 32776)) + list(range(65529, 65537))
    for n in sizes:
        data = base[1000:1000 + n]
        comp, _ = emu.deflate(data, level=1)
        assert zlib.decompress(comp, -15) == data, n
    for n in (65537, 65536 + 32768, 131071, 131073, 196609):
        for level in (1, 6):
            data = base[:n]
            comp, lens = emu.deflate(data, level=level)
            assert zlib.decompress(comp, -15) == data, (n, level)
            assert len(lens) == (n + 65535) // 65536
    # long matches that straddle the sub-block and chunk boundaries
    rep = (b"0123456789abcdefghijklmnopqrstuvwxyz" * 8000)[:150_000]
    comp, _ = emu.deflate(rep, level=6)
    assert zlib.decompress(comp, -15) == rep and len(comp) < 10_000


def test_emu_deflate_history_variant(emu, orc):
    """levels 6-9 (the kernel's history variant): the second unit of a chunk refers back into the first; with the DICT flag (the buffer
    is ONE stream) a chunk also refers back into the 32 KiB before it. Every stream must inflate to its input (zlib and the oracle),
    the output with history must not be larger than without, and distances never exceed 32768 (zlib would reject the stream)."""
    text = datagen.text_like(460_001, 23)
    for data in (text, datagen.mixed(300_000, 24), (datagen.text_like(40_000, 25) * 12)[:333_333], bytes(200_000), datagen.random_bytes(150_000, 26),
                 text[:32768], text[:32769], text[:65536], text[:65537], text[:98304 + 5]):
        sizes = {}
        for level, flags in ((3, 1), (6, 1), (6, 3), (9, 3)):
            comp, lens = emu.deflate(data, level=level, final=flags)
            assert zlib.decompress(comp, -15) == data, (len(data), level, flags)
            err, out, cons = orc.inflate(comp, len(data) + 8)
            assert err == 0 and out == data and cons == len(comp)
            sizes[(level, flags)] = len(comp)
        if len(data) > 100_000 and data[:64] != bytes(64):
            assert sizes[(6, 3)] <= sizes[(6, 1)] <= sizes[(3, 1)] * 1.002, sizes
    # a period longer than a unit: only the history can find it (distance 40 000 is out of reach, 30 000 is not)
    for period, reachable in ((30_000, True), (40_000, False)):
        blockdata = datagen.random_bytes(period, 27)
        data = (blockdata * 6)[:170_000]
        comp, _ = emu.deflate(data, level=6, final=3)
        assert zlib.decompress(comp, -15) == data
        assert (len(comp) < 0.5 * len(data)) == reachable, (period, len(comp))


def test_emu_deflate_runs_and_near_sources(emu, orc):
    """the M0 pass: runs and short periods take a source 1..4 positions back and only the anchors of a run keep their match; the
    traps of datagen.near_period_traps must not be mistaken for such sources; ratios on runs stay in zlib-1's neighbourhood"""
    traps = datagen.near_period_traps()
    cases = [traps, bytes(200_000), (b"abc" * 70_000)[:200_001], b"".join(bytes([i & 255]) * (5 + 13 * i % 700) for i in range(900)),
             b"".join(datagen.random_bytes(37, i) + bytes(3000 + 17 * i) for i in range(60)), b"".join((i // 7).to_bytes(4, "little") for i in range(60_000)),
             bytes(7) + b"x" + bytes(9) + b"y" * 11 + bytes(300), traps[:32768] + bytes(40_000) + traps[:20_000]]
    for data in cases:
        for level, flags in ((1, 1), (3, 1), (6, 1), (9, 3)):
            comp, _ = emu.deflate(data, level=level, final=flags)
            assert zlib.decompress(comp, -15) == data, (len(data), level, flags)
            err, out, cons = orc.inflate(comp, len(data) + 8)
            assert err == 0 and out == data and cons == len(comp)
    comp, _ = emu.deflate(bytes(1 << 20), level=1)
    assert len(comp) < (1 << 20) * 0.006  # 32 KiB of zeros: ~130 bytes (zlib level 1: 0.0044)


def test_emu_deflate_ratio_sane(emu):
    data = datagen.text_like(1 << 18, 9)
    comp, _ = emu.deflate(data, level=1)
    z = zlib.compress(data, 1)
    assert len(comp) < 1.1 * len(z)  # level-1 profile is in zlib level-1 territory


@pytest.mark.parametrize("n", [0, 1, 17, 512, 513, 70000])
def test_emu_crc(emu, n):
    data = datagen.random_bytes(n, n)
    for mis in (0, 3):
        v, segs = emu.crc32(data, 4096, mis)
        assert v == zlib.crc32(data)


def test_emu_inflate_windows(emu):
    data = datagen.mixed(150000, 6)
    for level in (0, 1, 9):
        co = zlib.compressobj(level, zlib.DEFLATED, -15)
        comp = co.compress(data) + co.flush()
        for iw, ow in ((0, 0), (3000, 66000)):
            st, out, cons, _ = emu.inflate(comp + b"\x55" * 9, len(data), iw, ow)
            assert st == 1 and out == data and cons == len(comp)
    st, _, _, _ = emu.inflate(comp[:1000], len(data))
    assert st == -5


def _raw(data, level, flush_every=0):
    co = zlib.compressobj(level, zlib.DEFLATED, -15)
    if not flush_every:
        return co.compress(data) + co.flush()
    parts = []
    for o in range(0, len(data), flush_every):  # Z_FULL_FLUSH forces block boundaries (and stored empty blocks) often
        parts.append(co.compress(data[o:o + flush_every]))
        parts.append(co.flush(zlib.Z_FULL_FLUSH))
    parts.append(co.flush())
    return b"".join(parts)


@pytest.mark.parametrize("level,flush", [(6, 0), (1, 0), (9, 20000), (6, 5000)])
def test_emu_inflate_speculative_rounds(emu, level, flush):
    """K6 on the emulator: guessed block starts, symbolic scan, chain proof, window resolve, emit -- bit-exact with zlib,
    and at least some segments really went through the speculative path (not only the serial fallback)."""
    data = datagen.text_like(500_000, seed=50 + level) + datagen.binary_records(60_000, seed=2) + datagen.text_like(200_000, seed=9)
    comp = _raw(data, level, flush)
    st, out, cons, stats = emu.inflate_spec(comp + b"\xaa" * 7, len(data), seg_bytes=8192, max_seg=64)
    assert st == 1 and cons == len(comp), (st, cons, len(comp), stats)
    assert out == data
    assert stats["discarded"] == 0
    assert stats["chain"] >= 4, stats  # several segments were decoded from guessed starts and proven
    # bounded compressed windows (the vtbl path's shape): each round sees only 64 KiB of input
    st, out, cons, stats2 = emu.inflate_spec(comp, len(data), seg_bytes=4096, max_seg=16, window=65536)
    assert st == 1 and out == data and cons == len(comp), stats2


def test_emu_inflate_speculative_hostile_inputs(emu):
    """stored blocks, fixed-Huffman blocks, tiny streams, truncation and corruption: the rounds may give up, never lie"""
    rnd = datagen.random_bytes(200_000, seed=4)          # level 6 on noise -> stored blocks only: no dynamic header to find
    comp = _raw(rnd, 6)
    st, out, cons, stats = emu.inflate_spec(comp, len(rnd), seg_bytes=4096, max_seg=32)
    assert st == 1 and out == rnd and cons == len(comp)
    small = b"abcabcabc" * 20                             # one fixed block
    comp = _raw(small, 6)
    st, out, cons, _ = emu.inflate_spec(comp, len(small), seg_bytes=4096, max_seg=8)
    assert st == 1 and out == small
    data = datagen.text_like(300_000, seed=77)
    comp = _raw(data, 6)
    st, out, _, _ = emu.inflate_spec(comp[:len(comp) // 2], len(data), seg_bytes=4096, max_seg=32)
    assert st == -5 and data.startswith(out[:1000])      # truncated: zlib's Z_BUF_ERROR
    bad = bytearray(comp)
    for k in range(len(bad) // 2, len(bad) // 2 + 40):
        bad[k] ^= 0x5A
    st, out, _, stats = emu.inflate_spec(bytes(bad), len(data) + 70000, seg_bytes=4096, max_seg=32)
    ref = zlib.decompressobj(-15)
    try:
        good = ref.decompress(bytes(bad))
        ok = ref.eof
    except zlib.error:
        good, ok = None, False
    if ok:
        assert st == 1 and out == good
    else:
        assert st < 0, (st, stats)


def test_emu_inflate_speculative_long_chain(emu):
    """more chain members than K6d has groups: the compose / link / resolve steps really fold several members per group"""
    data = datagen.text_like(700_000, seed=91)
    comp = _raw(data, 6, flush_every=2500)  # ~280 dynamic blocks, each followed by an empty stored block
    st, out, cons, stats = emu.inflate_spec(comp, len(data), seg_bytes=1024, max_seg=1024)
    assert st == 1 and out == data and cons == len(comp), stats
    assert stats["chain"] > 200 and stats["rounds"] <= 3 and stats["discarded"] == 0, stats


def test_emu_inflate_speculative_never_lies_under_corruption(emu):
    """seeded bit flips, byte runs and splices in a valid member: whatever the speculative rounds prove must be exactly
    what zlib produces -- same bytes if zlib accepts the damaged stream, an error if zlib rejects it, and never more
    clean output than zlib delivers before the error"""
    import random
    rng = random.Random(20260922)
    data = datagen.text_like(260_000, seed=5) + datagen.binary_records(40_000, seed=6)
    comp = _raw(data, 6, flush_every=0)
    for trial in range(12):
        bad = bytearray(comp)
        kind = trial % 3
        pos = rng.randrange(len(bad) // 8, len(bad) - 64)
        if kind == 0:
            bad[pos] ^= 1 << rng.randrange(8)
        elif kind == 1:
            for k in range(rng.randrange(2, 40)):
                bad[pos + k] = rng.randrange(256)
        else:  # splice: a piece of the stream copied over another place (valid-looking headers in the wrong spot)
            src = rng.randrange(0, len(bad) - 4096)
            bad[pos:pos + 2000] = comp[src:src + 2000]
        ref = zlib.decompressobj(-15)
        good = b""
        try:
            for o in range(0, len(bad), 4096):  # piecewise, so the bytes zlib delivered before an error are kept
                good += ref.decompress(bytes(bad[o:o + 4096]))
            ok = ref.eof
        except zlib.error:
            ok = False
        st, out, cons, stats = emu.inflate_spec(bytes(bad), len(data) + 100_000, seg_bytes=2048, max_seg=256)
        if ok:
            assert st == 1 and out == good, (trial, kind, pos, st, stats)
        else:
            assert st < 0, (trial, kind, pos, st, stats)
            n = min(len(out), len(good))
            assert out[:n] == good[:n], (trial, kind, pos)


def test_emu_degenerate_inputs_both_directions(emu):
    """SURVEY 8c's degenerate streams: all-'A' input gives zlib blocks of megabytes made of distance-1 self-overlapping
    runs (the overlap-aware copy, huge output per compressed byte); a 256-byte cycle and a 2-byte period exercise the
    short-period copy paths. Decode (serial and speculative) must match zlib; our encoder's output must decode."""
    data = b"A" * 3_000_000 + bytes(range(256)) * 2000 + b"AB" * 400_000 + b"\x00" * 70_000
    for level in (1, 6):
        comp = _raw(data, level)
        st, out, cons, _ = emu.inflate(comp, len(data))
        assert st == 1 and out == data and cons == len(comp)
        st, out, cons, stats = emu.inflate_spec(comp, len(data), seg_bytes=1024, max_seg=64)
        assert st == 1 and out == data and cons == len(comp), stats
        ours, _ = emu.deflate(data, level=level)
        assert zlib.decompress(ours, -15) == data and len(ours) < len(data) // 20


def test_emu_inflate_resumes_through_tiny_windows(emu):
    """the decoder is resumable at symbol granularity: input fed 33 bytes at a time, output space granted 258 / 300 bytes at a
    time (a match may have to be un-read and retried), stored / fixed / dynamic blocks alike"""
    data = datagen.mixed(30_000, 3) + b"Q" * 3000 + datagen.text_like(12_000, 4)
    for level in (0, 6):
        comp = _raw(data, level)
        for iw, ow in ((33, 0), (0, 258), (13, 300)):
            st, out, cons, _ = emu.inflate(comp, len(data), iw, ow)
            assert st == 1 and out == data and cons == len(comp), (level, iw, ow, st)
    tiny = _raw(b"abcabcabcabc", 6)  # one fixed-Huffman block
    for iw, ow in ((1, 0), (0, 1), (2, 3)):
        st, out, cons, _ = emu.inflate(tiny, 12, iw, ow)
        assert st == 1 and out == b"abcabcabcabc" and cons == len(tiny)


def test_emu_deflate_literal_heavy_blocks_take_the_split_path(emu):
    """low-entropy data without repeats: a 32 KiB sub-block is 32 768 literal tokens, more than the 24 576-entry token
    list holds, so it is coded as two blocks; still smaller than stored, and decodable"""
    import random
    rng = random.Random(5)
    nib = bytes(rng.randrange(16) for _ in range(200_000))
    mixed = nib[:70_000] + b"x" * 50_000 + nib[70_000:140_000]
    for data in (nib, mixed):
        for level in (1, 6):
            comp, _ = emu.deflate(data, level=level)
            assert zlib.decompress(comp, -15) == data
            assert len(comp) < 0.6 * len(data)


def test_emu_sha256_known_answers_and_every_padding_case(emu):
    """K7 against FIPS 180-4 known answers and hashlib: every length around the one-/two-block padding boundary (55, 56, 63, 64),
    aligned (16-byte loads) and unaligned (byte loads) starts, empty and multi-block messages, many messages per launch."""
    import hashlib
    assert emu.sha256([b"abc"])[0].hex() == "ba7816bf8f01cfea414140de5dae2223b00361a396177a9cb410ff61f20015ad"
    assert emu.sha256([b""])[0].hex() == "e3b0c44298fc1c149afbf4c8996fb92427ae41e4649b934ca495991b7852b855"
    assert emu.sha256([b"abcdbcdecdefdefgefghfghighijhijkijkljklmklmnlmnomnopnopq"])[0].hex() == \
        "248d6a61d20638b8e5c026930c3e6039a33ce45964ff2167f6ecedd419db06c1"
    base = datagen.random_bytes(70_000, 5)
    msgs = [base[:n] for n in list(range(0, 130)) + [255, 256, 4095, 4096, 65535, 65536, 70_000]]
    for shift in (0, 1, 7):
        got = emu.sha256(msgs, align=16, shift=shift)
        for m, g in zip(msgs, got):
            assert g == hashlib.sha256(m).digest(), (len(m), shift)
    many = [datagen.text_like(1000 + 37 * i, seed=i)[: 1000 + 37 * i] for i in range(300)]
    for m, g in zip(many, emu.sha256(many)):
        assert g == hashlib.sha256(m).digest()


def test_emu_fuzzers_short(emu):
    """a short, seeded turn of the two randomized differential runs (tests/emu/fuzz_kernels.py: every kernel against zlib;
    tests/emu/fuzz_deflate.py: the deflate kernel at random levels, with and without the one-stream flag, over runs and periods).
    Run them longer by hand after touching a kernel: the period-3 bug of round 2 took 277 iterations to show."""
    import sys
    emu_dir = os.path.join(HERE, "emu")
    for script, seed, secs in (("fuzz_kernels.py", 20260923, 15), ("fuzz_deflate.py", 7, 10)):
        r = subprocess.run([sys.executable, os.path.join(emu_dir, script), str(seed), str(secs)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
        assert r.returncode == 0 and b"fails 0" in r.stdout, r.stdout[-2000:]


def test_emu_kernels_under_sanitizers():
    """the same kernel sources built with AddressSanitizer + UBSan: out-of-bounds global accesses, shifts by >= 32,
    signed overflow ... in the deflate, CRC, inflate and K6 kernels abort the run"""
    here = os.path.dirname(os.path.abspath(__file__))
    emu_dir = os.path.join(here, "emu")
    r = subprocess.run(["make", "-s", "-C", emu_dir, "libmzemu_san.so"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    asan = subprocess.run(["/usr/bin/gcc", "-print-file-name=libasan.so"], stdout=subprocess.PIPE).stdout.decode().strip()
    if r.returncode != 0 or not os.path.isabs(asan) or not os.path.exists(asan):
        pytest.skip("no sanitizer runtime in this toolchain")
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0:halt_on_error=1")
    import sys
    r = subprocess.run([sys.executable, os.path.join(emu_dir, "san_run.py")], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=900)
    assert r.returncode == 0 and b"sanitized run ok" in r.stdout, r.stdout[-3000:]
