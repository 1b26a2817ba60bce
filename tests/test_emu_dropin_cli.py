"""The drop-in proof without a GPU: the REFERENCE's `minigzip` / `minizip` programs and container code, compiled where they
lie with `mz_strm_zlib.c` left out and its names aliased to `mz_stream_cuda_*` (oracle/Makefile `emu` targets), linked against
the product built on the CPU execution-model emulator (tests/emu/libmz_strm_emu.so = real host C + real API shim + kernel
sources). Same checks as tests/test_gpu_dropin_cli.py on smaller inputs: both directions against the unmodified reference
builds and Python's gzip / zipfile; the batch zip writer / extractor on the raw-entry seam."""
import gzip
import json
import os
import subprocess
import zipfile
import zlib

import pytest

import datagen

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REFDIR = os.path.join(ROOT, "oracle", "_ref")
ENV = dict(os.environ, MZ_CUDA_BATCH_KB="512", MZ_CUDA_SPEC_SEG_KB="4")


@pytest.fixture(scope="module")
def emu_bins(built):
    if not os.path.exists("/root/reference/mz_strm_zlib.c"):
        pytest.skip("needs the reference sources at build time")
    r = subprocess.run(["make", "-s", "-C", os.path.join(HERE, "emu"), "libmz_strm_emu.so"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-3000:]
    r = subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "emu"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-3000:]
    return {n: os.path.join(REFDIR, n) for n in ("minigzip_emu", "minizip_emu", "zipbatch_emu", "minigzip_ref", "minizip_ref")}


def _run(args, cwd, ok=True):
    r = subprocess.run(args, cwd=cwd, env=ENV, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=900)
    if ok:
        assert r.returncode == 0, (args, r.returncode, r.stdout[-600:], r.stderr[-600:])
    return r


def _corpus():
    return {"text.txt": datagen.text_like(900_000, seed=21), "records.bin": datagen.binary_records(300_000, seed=22),
            "random.bin": datagen.random_bytes(100_000, seed=23), "tiny.txt": b"hello, hello, hello\n", "empty.bin": b""}


@pytest.mark.parametrize("level", [1, 6])
def test_minigzip_both_directions(emu_bins, tmp_path, level):
    data = datagen.text_like(1_200_000, seed=30 + level) + datagen.random_bytes(50_000, seed=3)
    (tmp_path / "doc.txt").write_bytes(data)
    r = _run([emu_bins["minigzip_emu"], f"-{level}", "doc.txt"], tmp_path)
    assert b"Operation completed successfully" in r.stdout
    gz = (tmp_path / "doc.txt.gz").read_bytes()
    assert gzip.decompress(gz) == data and int.from_bytes(gz[-8:-4], "little") == zlib.crc32(data)
    _run([emu_bins["minigzip_ref"], "-x", "-d", "x", "doc.txt.gz"], tmp_path)
    assert (tmp_path / "x" / "doc.txt").read_bytes() == data
    (tmp_path / "a.bin").write_bytes(data)
    _run([emu_bins["minigzip_ref"], "-9", "a.bin"], tmp_path)
    _run([emu_bins["minigzip_emu"], "-x", "-d", "o1", "a.bin.gz"], tmp_path)
    assert (tmp_path / "o1" / "a.bin").read_bytes() == data
    bad = bytearray((tmp_path / "a.bin.gz").read_bytes())
    bad[len(bad) // 2] ^= 0x10
    (tmp_path / "c.bin.gz").write_bytes(bytes(bad))
    r = _run([emu_bins["minigzip_emu"], "-x", "-d", "o3", "c.bin.gz"], tmp_path, ok=False)
    assert r.returncode != 0 and b"Error" in r.stdout


def test_minizip_both_directions(emu_bins, tmp_path):
    files = _corpus()
    for name, blob in files.items():
        (tmp_path / name).write_bytes(blob)
    _run([emu_bins["minizip_emu"], "-o", "-6", "a.zip"] + sorted(files), tmp_path)
    with zipfile.ZipFile(tmp_path / "a.zip") as z:
        assert z.testzip() is None and sorted(z.namelist()) == sorted(files)
        for name, blob in files.items():
            assert z.read(name) == blob and z.getinfo(name).CRC == zlib.crc32(blob)
    _run([emu_bins["minizip_ref"], "-x", "-o", "-d", "out", "a.zip"], tmp_path)  # the unmodified reference extracts and CRC-checks
    _run([emu_bins["minizip_ref"], "-o", "-9", "r.zip"] + sorted(files), tmp_path)
    _run([emu_bins["minizip_emu"], "-x", "-o", "-d", "o1", "r.zip"], tmp_path)   # ... and we extract what it wrote
    for name, blob in files.items():
        assert (tmp_path / "out" / name).read_bytes() == blob and (tmp_path / "o1" / name).read_bytes() == blob


def test_zip_batch_writer_and_extractor(emu_bins, tmp_path):
    exe = emu_bins["zipbatch_emu"]
    n, esz = 260, 30_000
    (tmp_path / "dump").mkdir()
    st = json.loads(_run([exe, "c.zip", str(n), str(esz), "6", "cuda", "dump", "13"], tmp_path).stdout.decode().strip().splitlines()[-1])
    assert st["err"] == 0 and st["close_err"] == 0 and st["entries"] == n
    _run([exe, "r.zip", str(n), str(esz), "6", "ref"], tmp_path)
    with zipfile.ZipFile(tmp_path / "c.zip") as zc, zipfile.ZipFile(tmp_path / "r.zip") as zr:
        assert zc.testzip() is None and zc.namelist() == zr.namelist() == ["e/%06d" % i for i in range(n)]
        for i in range(n):
            a, b = zc.getinfo("e/%06d" % i), zr.getinfo("e/%06d" % i)
            assert a.CRC == b.CRC and a.file_size == b.file_size
        for i in range(0, n, 13):
            assert zc.read("e/%06d" % i) == (tmp_path / "dump" / ("%06d" % i)).read_bytes()
    _run([emu_bins["minizip_ref"], "-x", "-o", "-d", "out", "c.zip"], tmp_path)
    for name in ("c.zip", "r.zip"):
        got = json.loads(_run([exe, name, str(n), str(esz), "0", "extract"], tmp_path).stdout.decode().strip().splitlines()[-1])
        want = json.loads(_run([exe, name, str(n), str(esz), "0", "extract_ref"], tmp_path).stdout.decode().strip().splitlines()[-1])
        assert got["err"] == 0 and got["entries"] == want["entries"] == n and got["bytes_out"] == want["bytes_out"] and got["mismatches"] == 0
    raw = bytearray((tmp_path / "r.zip").read_bytes())
    with zipfile.ZipFile(tmp_path / "r.zip") as z:
        info = z.getinfo("e/000100")
    raw[info.header_offset + 30 + len(info.filename) + info.compress_size // 2] ^= 0x40
    (tmp_path / "bad.zip").write_bytes(bytes(raw))
    r = _run([exe, "bad.zip", str(n), str(esz), "0", "extract"], tmp_path, ok=False)
    assert r.returncode != 0 and json.loads(r.stdout.decode().strip().splitlines()[-1])["err"] in (-3, -105)


def _sha_roundtrip(exe, refc, tmp_path, run):
    """scope row f3: per-entry SHA-256 computed by K7 and stored in MZ_ZIP_EXTENSION_HASH exactly as the reference's writer stores
    it (mz_zip_rw.c:1398-1408); checked three ways -- hashlib over every entry, the reference built WITH its crypto provider
    (oracle/_ref/minizip_refc verifies the hash while extracting, mz_zip_rw.c:430-450), and the product's own batch extractor --
    in both directions (the reference's hashes are verified by K7 too), and a damaged digest must be refused by both."""
    import hashlib
    n, esz = 120, 20_000
    st = json.loads(run([exe, "h.zip", str(n), str(esz), "6", "cuda_sha"], tmp_path).stdout.decode().strip().splitlines()[-1])
    assert st["err"] == 0 and st["close_err"] == 0 and st["entries"] == n
    with zipfile.ZipFile(tmp_path / "h.zip") as z:
        assert z.testzip() is None
        for info in z.infolist():
            x = info.extra
            assert x[:8] == bytes([0x51, 0x1a, 36, 0, 23, 0, 32, 0]), x.hex()
            assert x[8:40] == hashlib.sha256(z.read(info)).digest(), info.filename
        victim = z.getinfo("e/000050")
    run([refc, "-x", "-o", "-d", "out_h", "h.zip"], tmp_path)                      # the reference verifies every hash
    got = json.loads(run([exe, "h.zip", str(n), str(esz), "0", "extract"], tmp_path).stdout.decode().strip().splitlines()[-1])
    assert got["err"] == 0 and got["entries"] == n and got["mismatches"] == 0
    # the other direction: an archive whose hashes the REFERENCE computed (OpenSSL) is verified by K7
    (tmp_path / "src").mkdir()
    for k, v in _corpus().items():
        (tmp_path / "src" / k).write_bytes(v)
    run([refc, "-o", "-6", "../r_h.zip"] + sorted(_corpus()), tmp_path / "src")
    with zipfile.ZipFile(tmp_path / "r_h.zip") as z:
        assert all(i.extra[:2] == b"\x51\x1a" for i in z.infolist())
    got = json.loads(run([exe, "r_h.zip", "0", str(1 << 20), "0", "extract"], tmp_path).stdout.decode().strip().splitlines()[-1])
    assert got["err"] == 0 and got["entries"] == len(_corpus())
    # one flipped digest bit in the central directory: refused by the reference reader and by the batch extractor (MZ_CRC_ERROR)
    raw = bytearray((tmp_path / "h.zip").read_bytes())
    cd = raw.rfind(b"PK\x01\x02" + b"", 0)
    pos = 0
    while True:  # walk the central directory to the victim's record
        pos = raw.find(b"PK\x01\x02", pos)
        assert pos >= 0
        nlen = int.from_bytes(raw[pos + 28:pos + 30], "little")
        if raw[pos + 46:pos + 46 + nlen] == victim.filename.encode():
            break
        pos += 46
    raw[pos + 46 + nlen + 8 + 5] ^= 0x01
    (tmp_path / "bad_h.zip").write_bytes(bytes(raw))
    r = run([exe, "bad_h.zip", str(n), str(esz), "0", "extract"], tmp_path, ok=False)
    assert r.returncode != 0 and json.loads(r.stdout.decode().strip().splitlines()[-1])["err"] == -105
    r = run([refc, "-x", "-o", "-d", "out_bad", "bad_h.zip"], tmp_path, ok=False)
    assert r.returncode != 0 or b"rror" in r.stdout + r.stderr


def test_zip_batch_sha256_extrafield(emu_bins, tmp_path):
    refc = os.path.join(REFDIR, "minizip_refc")
    if not os.path.exists(refc):
        pytest.skip("oracle/_ref/minizip_refc not built (no OpenSSL headers)")
    _sha_roundtrip(emu_bins["zipbatch_emu"], refc, tmp_path, _run)


def _native_archive_checks(exe, ref_bins, tmp_path, run, n=260, esz=30_000):
    """mz_zip_cuda_write_archive: the product writes local headers, streams, central directory and end records itself; the
    archive must be what the reference's container would have produced as far as any reader can tell: same names / CRCs / sizes
    as the reference-written archive, extractable by the unmodified reference CLI and by CPython's zipfile, readable by the
    batch extractor."""
    (tmp_path / "dumpn").mkdir()
    st = json.loads(run([exe, "n.zip", str(n), str(esz), "6", "native", "dumpn", "17"], tmp_path).stdout.decode().strip().splitlines()[-1])
    assert st["err"] == 0 and st["entries"] == n
    run([exe, "rn.zip", str(n), str(esz), "6", "ref"], tmp_path)
    with zipfile.ZipFile(tmp_path / "n.zip") as zn, zipfile.ZipFile(tmp_path / "rn.zip") as zr:
        assert zn.testzip() is None and zn.namelist() == zr.namelist()
        for a, b in zip(zn.infolist(), zr.infolist()):
            assert (a.CRC, a.file_size, a.compress_type, a.date_time) == (b.CRC, b.file_size, b.compress_type, b.date_time), a.filename
            assert a.flag_bits & 0x808 == 0x800 and a.external_attr >> 16 == 0o100644
        for i in range(0, n, 17):
            assert zn.read("e/%06d" % i) == (tmp_path / "dumpn" / ("%06d" % i)).read_bytes()
    run([ref_bins["minizip_ref"], "-x", "-o", "-d", "outn", "n.zip"], tmp_path)
    assert (tmp_path / "outn" / "e" / "000034").read_bytes() == (tmp_path / "dumpn" / "000034").read_bytes()
    got = json.loads(run([exe, "n.zip", str(n), str(esz), "0", "extract"], tmp_path).stdout.decode().strip().splitlines()[-1])
    want = json.loads(run([exe, "n.zip", str(n), str(esz), "0", "extract_ref"], tmp_path).stdout.decode().strip().splitlines()[-1])
    assert got["err"] == 0 and want["err"] == 0 and got["entries"] == want["entries"] == n and got["bytes_out"] == want["bytes_out"]


def test_native_archive_writer(emu_bins, tmp_path):
    _native_archive_checks(emu_bins["zipbatch_emu"], emu_bins, tmp_path, _run)
    # 1 MiB rounds: eight of them, four in preparation at a time on their own worker threads
    st = json.loads(_run(["env", "MZ_CUDA_ZIP_ROUND_MB=1", emu_bins["zipbatch_emu"], "n8.zip", "260", "30000", "6", "native"], tmp_path).stdout.decode().strip().splitlines()[-1])
    assert st["err"] == 0 and st["rounds"] >= 8
    with zipfile.ZipFile(tmp_path / "n8.zip") as z8, zipfile.ZipFile(tmp_path / "n.zip") as z1:
        assert z8.testzip() is None and [(i.filename, i.CRC, i.file_size) for i in z8.infolist()] == [(i.filename, i.CRC, i.file_size) for i in z1.infolist()]
