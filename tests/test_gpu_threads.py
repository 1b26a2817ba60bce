"""Several host threads, one device: the reference is single-threaded per stream object but distinct objects are
independent (SURVEY 8b "Threading"); the CUDA backend shares a process-wide context, a workspace pool, work counters and
CRC scratch, so concurrent streams must not disturb each other."""
import ctypes as C
import threading
import zlib

import pytest

import datagen

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env(built):
    import cuharness
    p = cuharness.pkg()
    lib = p.load()
    assert lib.mz_cuda_init() == 0
    return p, lib


def test_concurrent_streams_and_crc_calls(env):
    import cuharness
    p, lib = env
    errors = []

    def worker(k):
        try:
            tl = cuharness.TestLib()  # own ctypes handle per thread
            data = datagen.mixed(2_500_000 + 111_111 * k, seed=100 + k) + datagen.text_like(1_000_000, seed=k)
            for rep in range(3):
                level = (1, 6, 9)[(k + rep) % 3]
                wbits = (31, -15, 15)[(k + rep) % 3]
                comp, info = tl.compress(lib.mz_stream_cuda_create, data, level=level, window_bits=wbits, write_size=65536 + 17 * k)
                assert info["close"] == 0 and info["total_in"] == len(data)
                assert zlib.decompress(comp, wbits) == data
                out, rinfo = tl.decompress(lib.mz_stream_cuda_create, comp, len(data), window_bits=wbits, read_size=50_000 + k)
                assert rinfo["read"] == len(data) and out == data and rinfo["total_in"] == len(comp)
                buf = C.create_string_buffer(data, len(data))
                assert lib.mz_crypt_crc32_update(0, buf, len(data)) == zlib.crc32(data)  # > 1 MiB: device path, shared scratch
        except Exception as e:  # noqa: BLE001 -- collected and re-raised in the main thread
            errors.append((k, repr(e)))

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(6)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_crc_symbol_hammered_from_eight_threads(env):
    """mz_crypt_crc32_update is a pure, re-entrant function in the reference (mz_crypt.c:35-92): eight threads call the
    replacement at once with different buffers between 1 and 4 MiB (all on the device path, which shares per-device staging),
    chained in uneven pieces, and every single result is checked. Also below the threshold (host slice-by-16 path)."""
    p, lib = env
    import numpy as np
    errors = []

    def worker(k):
        try:
            rng = np.random.default_rng(1000 + k)
            for rep in range(12):
                n = int(rng.integers(1 << 20, 4 << 20))
                data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
                buf = C.create_string_buffer(data, n)
                want = zlib.crc32(data)
                got = lib.mz_crypt_crc32_update(0, buf, n)
                assert got == want, ("whole", k, rep, n, hex(got), hex(want))
                cut = int(rng.integers(1 << 20, n)) if n > (1 << 20) + 1 else n
                part = lib.mz_crypt_crc32_update(0, buf, cut)           # device path
                tail = C.create_string_buffer(data[cut:], n - cut)     # may be short: host path
                got2 = lib.mz_crypt_crc32_update(part, tail, n - cut)
                assert got2 == want, ("chained", k, rep, n, cut)
                small = data[: int(rng.integers(1, 70000))]
                sb = C.create_string_buffer(small, len(small))
                assert lib.mz_crypt_crc32_update(0, sb, len(small)) == zlib.crc32(small)
        except Exception as e:  # noqa: BLE001
            errors.append((k, repr(e)))

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
