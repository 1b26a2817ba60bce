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
